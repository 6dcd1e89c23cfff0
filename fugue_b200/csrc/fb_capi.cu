// C-ABI plumbing shared by all kernels: error strings, device discovery.
#include <stdarg.h>

#include <mutex>

#include "fb_common.cuh"

static thread_local char g_err[1024] = "";

void fb_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int fb_sm_count(int dev) {
  static int cache[64];
  static std::mutex mu;
  if (dev < 0 || dev >= 64) return 148;
  std::lock_guard<std::mutex> lock(mu);
  if (cache[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    cache[dev] = v;
  }
  return cache[dev];
}

extern "C" {

int fb_abi_version(void) { return FB_ABI_VERSION; }

const char* fb_last_error(void) { return g_err; }

int fb_device_info(int dev, int* sm_count, size_t* total_mem, int* cc_major, int* cc_minor) {
  cudaDeviceProp p;
  FB_CUDA(cudaGetDeviceProperties(&p, dev));
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (total_mem) *total_mem = p.totalGlobalMem;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  return 0;
}

uint32_t fb_debug_fastmod_host(uint64_t hash, uint32_t num_partitions) {
  FbDiv dv = fb_make_div(num_partitions);
  return fb_fastmod(hash, dv);
}

}  // extern "C"
