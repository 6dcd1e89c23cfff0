// Design probe for K6 (B200): multi-pass filtered aggregation.  Pass p streams ALL keys, keeps the rows whose hash
// falls into slice p of P, and aggregates them into ONE table slice that stays L2-resident (P x slice = whole table).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a groupby_probe.cu -o groupby_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t h) {
  h ^= h >> 30; h *= 0xBF58476D1CE4E5B9ULL; h ^= h >> 27; h *= 0x94D049BB133111EBULL; h ^= h >> 31; return h;
}
__global__ void gen(uint64_t* keys, double* vals, int64_t n, uint64_t nk) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    keys[i] = (mix((uint64_t)i) % nk) * 0x9E3779B97F4A7C15ULL % (1ULL << 62);
    vals[i] = (double)(i & 1023) * 0.5;
  }
}
__global__ void init(uint64_t* t, int64_t words) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < words; i += (int64_t)gridDim.x * blockDim.x)
    t[i] = (i & 3) == 0 ? ~0ULL : 0ULL;
}
// one warp: scan 32 x UNROLL rows, queue the matching (key, row) pairs in shared memory, and whenever 32 are queued
// let ALL lanes do one find-or-insert + update each (the slow, latency-bound part runs with full warps)
constexpr int kWarpsPerCta = 8;
template <int UNROLL>
__global__ void __launch_bounds__(kWarpsPerCta * 32) pass_kernel(const uint64_t* __restrict__ keys, const double* __restrict__ vals,
                                                                int64_t n, uint64_t* __restrict__ table, uint64_t slice_mask,
                                                                uint32_t pass, uint32_t pshift, int* overflow) {
  constexpr int QCAP = 32 * (UNROLL + 1);
  __shared__ uint64_t q_key[kWarpsPerCta][QCAP];
  __shared__ uint32_t q_row[kWarpsPerCta][QCAP];
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1;
  uint64_t* qk = q_key[warp];
  uint32_t* qr = q_row[warp];
  int qn = 0;
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerCta;
  const int64_t gw = (int64_t)blockIdx.x * kWarpsPerCta + warp;
  // contiguous row range per warp, in steps of 32 * UNROLL rows
  const int64_t steps = (n + 32 * UNROLL - 1) / (32 * UNROLL);
  const int64_t s0 = steps * gw / nwarps, s1 = steps * (gw + 1) / nwarps;
  auto process = [&](uint64_t key, uint32_t row, bool active) {
    if (active) {
      const uint64_t h = mix(key);
      uint64_t s = (h >> 8) & slice_mask;
      const double v = __ldcs(vals + row);
      for (int probe = 0; probe < 4096; ++probe) {
        uint64_t* slot = table + s * 4;
        uint64_t cur = *(volatile uint64_t*)slot;
        if (cur == ~0ULL) {
          cur = atomicCAS((unsigned long long*)slot, ~0ULL, (unsigned long long)key);
          if (cur == ~0ULL) cur = key;
        }
        if (cur == key) {
          atomicAdd((double*)(slot + 1), v);
          atomicAdd((unsigned long long*)(slot + 2), 1ULL);
          break;
        }
        s = (s + 1) & slice_mask;
        if (probe == 4095) *overflow = 1;
      }
    }
  };
  for (int64_t st = s0; st < s1; ++st) {
    const int64_t base = st * (32 * UNROLL) + lane;
    uint64_t k[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int64_t i = base + u * 32;
      k[u] = i < n ? __ldcs((const unsigned long long*)keys + i) : 0;
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int64_t i = base + u * 32;
      const bool hit = i < n && (uint32_t)(mix(k[u]) >> pshift) == pass;
      const unsigned m = __ballot_sync(0xFFFFFFFFu, hit);
      if (hit) {
        const int pos = qn + __popc(m & lt);
        qk[pos] = k[u];
        qr[pos] = (uint32_t)i;
      }
      qn += __popc(m);
    }
    __syncwarp();
    while (qn >= 32) {
      qn -= 32;
      process(qk[qn + lane], qr[qn + lane], true);
      __syncwarp();
    }
  }
  __syncwarp();
  if (qn > 0) process(lane < (unsigned)qn ? qk[lane] : 0, lane < (unsigned)qn ? qr[lane] : 0, lane < (unsigned)qn);
}
__global__ void extract(const uint64_t* __restrict__ t, int64_t slots, uint64_t* ok, double* os, uint64_t* oc, unsigned long long* cursor) {
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < slots + 31; s += (int64_t)gridDim.x * blockDim.x) {
    const bool used = s < slots && t[s * 4] != ~0ULL;
    const unsigned m = __ballot_sync(0xFFFFFFFFu, used);
    if (m == 0) continue;
    unsigned long long b = 0;
    const int lead = __ffs(m) - 1;
    if ((threadIdx.x & 31) == lead) b = atomicAdd(cursor, (unsigned long long)__popc(m));
    b = __shfl_sync(0xFFFFFFFFu, b, lead);
    if (used) {
      const int64_t o = b + __popc(m & ((1u << (threadIdx.x & 31)) - 1));
      ok[o] = t[s * 4]; os[o] = ((const double*)t)[s * 4 + 1]; oc[o] = t[s * 4 + 2];
    }
  }
}
int main() {
  const int64_t n = 125000000; const uint64_t nk = 10000000;
  uint64_t *keys, *table, *ok, *oc; double *vals, *os; int* ovf; unsigned long long* cursor;
  CK(cudaMalloc(&keys, n * 8)); CK(cudaMalloc(&vals, n * 8));
  CK(cudaMalloc(&ok, 16000000 * 8)); CK(cudaMalloc(&os, 16000000 * 8)); CK(cudaMalloc(&oc, 16000000 * 8));
  CK(cudaMalloc(&ovf, 4)); CK(cudaMalloc(&cursor, 8));
  gen<<<148 * 8, 256>>>(keys, vals, n, nk);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float ms;
  for (int P : {4, 8, 16}) {
    const int64_t total_slots = 1 << 24;            // load factor 0.6 for 10 M groups
    const int64_t slots = total_slots / P;          // P = 8: 2 M slots x 32 B = 64 MB
    int pshift = 64; for (int q = P; q > 1; q >>= 1) --pshift;
    CK(cudaMalloc(&table, slots * 32));
    for (int grid_mult : {8, 4}) {
      for (int rep = 0; rep < 2; ++rep) {
        CK(cudaMemset(ovf, 0, 4)); CK(cudaMemset(cursor, 0, 8));
        cudaEventRecord(e0);
        for (int p = 0; p < P; ++p) {
          init<<<148 * 4, 256>>>(table, slots * 4);
          pass_kernel<8><<<148 * grid_mult, kWarpsPerCta * 32>>>(keys, vals, n, table, (uint64_t)slots - 1, (uint32_t)p, (uint32_t)pshift, ovf);
          extract<<<148 * 4, 256>>>(table, slots, ok, os, oc, cursor);
        }
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
      }
      int h_ovf; unsigned long long h_groups;
      cudaMemcpy(&h_ovf, ovf, 4, cudaMemcpyDeviceToHost); cudaMemcpy(&h_groups, cursor, 8, cudaMemcpyDeviceToHost);
      printf("P=%2d slice=%4lld MB grid=148x%d: %7.3f ms total (%6.1f G rows/s)  groups=%llu overflow=%d\n", P,
             (long long)(slots * 32 >> 20), grid_mult, ms, n / ms / 1e6, h_groups, h_ovf);
    }
    cudaFree(table);
  }
  return 0;
}
