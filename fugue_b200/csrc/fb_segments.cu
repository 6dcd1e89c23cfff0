// Segment copy: out[dst_off[s] .. +len[s]) = in[src_off[s] .. +len[s]) for every column.
// Used after the multi-GPU exchange to turn the received (source, partition) runs into
// (partition, source) order so that every owned partition is contiguous;
// the reference has no counterpart (its shuffles are Dask/Spark/Ray's, fugue_dask/_utils.py:124-130).
// Pure HBM copy: 2 x width bytes per row per column.
#include "fb_common.cuh"

namespace {

template <typename T>
__device__ __forceinline__ void copy_run(const T* __restrict__ src, T* __restrict__ dst, int64_t n) {
  const int64_t step = (int64_t)blockDim.x * 4;
  int64_t i = threadIdx.x;
  for (; i + 3 * (int64_t)blockDim.x < n; i += step) {  // 4 independent loads in flight per thread
    T a = src[i], b = src[i + blockDim.x], c = src[i + 2 * blockDim.x], d = src[i + 3 * blockDim.x];
    dst[i] = a; dst[i + blockDim.x] = b; dst[i + 2 * blockDim.x] = c; dst[i + 3 * blockDim.x] = d;
  }
  for (; i < n; i += blockDim.x) dst[i] = src[i];
}

__global__ void __launch_bounds__(256)
fb_copy_segments_kernel(const void* const* __restrict__ src_cols, void* const* __restrict__ dst_cols,
                        const int32_t* __restrict__ widths, const int64_t* __restrict__ src_off,
                        const int64_t* __restrict__ dst_off, const int64_t* __restrict__ len, int nseg,
                        int rows_per_cta) {
  const int c = blockIdx.y;
  const int w = widths[c];
  const uint8_t* s = (const uint8_t*)src_cols[c];
  uint8_t* d = (uint8_t*)dst_cols[c];
  // blockIdx.x enumerates (segment, piece) pairs: pieces of rows_per_cta rows, found by a walk
  // over the segment table is avoided by launching ceil(len/rows_per_cta) pieces per segment on
  // the host side via a piece table; here: one CTA per segment with an inner loop.
  for (int sgi = blockIdx.x; sgi < nseg; sgi += gridDim.x) {
    const int64_t n = len[sgi];
    if (n <= 0) continue;
    const int64_t so = src_off[sgi], dof = dst_off[sgi];
    switch (w) {
      case 8: copy_run((const uint64_t*)s + so, (uint64_t*)d + dof, n); break;
      case 4: copy_run((const uint32_t*)s + so, (uint32_t*)d + dof, n); break;
      case 2: copy_run((const uint16_t*)s + so, (uint16_t*)d + dof, n); break;
      default: copy_run(s + so, d + dof, n); break;
    }
  }
  (void)rows_per_cta;
}

}  // namespace

extern "C" int fb_copy_segments(int dev, void* stream, int ncols, const void* const* d_src_cols,
                                void* const* d_dst_cols, const int32_t* d_widths, int nseg,
                                const int64_t* d_src_off, const int64_t* d_dst_off,
                                const int64_t* d_len) {
  FB_CHECK(ncols >= 0 && nseg >= 0, "negative count");
  if (ncols == 0 || nseg == 0) return 0;
  FbDeviceGuard guard(dev);
  FB_CHECK(guard.ok, "cannot select device %d", dev);
  int gx = nseg < 148 * 8 ? nseg : 148 * 8;
  dim3 grid((unsigned)gx, (unsigned)ncols);
  fb_copy_segments_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_src_cols, d_dst_cols, d_widths, d_src_off,
                                                                 d_dst_off, d_len, nseg, 0);
  FB_CUDA(cudaGetLastError());
  return 0;
}
