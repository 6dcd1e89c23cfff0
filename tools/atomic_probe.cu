// Micro-benchmarks behind the K6 / K7 designs (B200): how many hash-table updates per second do the
// different memories sustain?  nvcc -O3 -gencode arch=compute_100a,code=sm_100a atomic_probe.cu -o atomic_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t h) {
  h ^= h >> 30; h *= 0xBF58476D1CE4E5B9ULL; h ^= h >> 27; h *= 0x94D049BB133111EBULL; h ^= h >> 31; return h;
}

// mode 0: RED f64 | 1: RED f64 + RED u64 same sector | 2: key load + compare + 2 RED | 3: 1 RED u64 only
// 4: key load only (probe) | 5: 16-byte load (join probe)
template <int MODE>
__global__ void l2_kernel(uint64_t* __restrict__ table, int64_t region_slots, int64_t rows_per_region, int64_t nrows,
                          uint64_t* sink) {
  uint64_t acc = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nrows; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t region = i / rows_per_region;
    const uint64_t h = mix((uint64_t)i * 0x9E3779B97F4A7C15ULL);
    uint64_t* slot = table + (region * region_slots + (int64_t)(h % (uint64_t)region_slots)) * 4;
    if (MODE == 0) atomicAdd((double*)(slot + 1), 1.0);
    if (MODE == 1) { atomicAdd((double*)(slot + 1), 1.0); atomicAdd((unsigned long long*)(slot + 2), 1ULL); }
    if (MODE == 2) {
      const uint64_t k = *(volatile uint64_t*)slot;
      if (k != 12345) { atomicAdd((double*)(slot + 1), 1.0); atomicAdd((unsigned long long*)(slot + 2), 1ULL); }
    }
    if (MODE == 3) atomicAdd((unsigned long long*)(slot + 2), 1ULL);
    if (MODE == 4) acc += *(volatile uint64_t*)slot;
    if (MODE == 5) { unsigned long long x, y; asm volatile("ld.global.relaxed.gpu.v2.u64 {%0, %1}, [%2];" : "=l"(x), "=l"(y) : "l"(slot)); acc += x + y; }
  }
  if (acc == 0x1234567) *sink = acc;
}

// shared-memory table: each CTA owns `slots` 20-byte logical slots (key u64, sum f64, count u32 in three arrays)
// mode 0: atomicAdd(double) + atomicAdd(u32) | 1: u32 count only | 2: key CAS-probe + both
template <int MODE>
__global__ void smem_kernel(int slots, int64_t nrows, uint64_t* sink) {
  extern __shared__ uint64_t sm[];
  uint64_t* keys = sm;
  double* sums = (double*)(sm + slots);
  uint32_t* cnts = (uint32_t*)(sm + 2 * slots);
  for (int i = threadIdx.x; i < slots; i += blockDim.x) { keys[i] = ~0ULL; sums[i] = 0; cnts[i] = 0; }
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nrows; i += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t h = mix((uint64_t)i * 0x9E3779B97F4A7C15ULL);
    const uint64_t key = h % (uint64_t)(slots / 2);  // load factor 0.5
    int s = (int)(mix(key) % (uint64_t)slots);
    if (MODE == 2) {
      while (true) {
        const uint64_t cur = keys[s];
        if (cur == key) break;
        if (cur == ~0ULL) {
          const uint64_t old = atomicCAS((unsigned long long*)&keys[s], ~0ULL, (unsigned long long)key);
          if (old == ~0ULL || old == key) break;
        }
        s = s + 1 == slots ? 0 : s + 1;
      }
    }
    if (MODE != 1) atomicAdd(&sums[s], 1.0);
    atomicAdd(&cnts[s], 1u);
  }
  __syncthreads();
  if (threadIdx.x == 0 && sums[0] == 1234567.0) *sink = 1;
}

int main() {
  const int64_t nrows = 125000000;
  const int64_t region_slots = 1 << 16;          // 65536 slots x 32 B = 2 MB per region (39K groups at load 0.6)
  const int nregions = 256;
  uint64_t *table, *sink;
  CK(cudaMalloc(&table, (size_t)nregions * region_slots * 32));
  CK(cudaMalloc(&sink, 8));
  CK(cudaMemset(table, 0, (size_t)nregions * region_slots * 32));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  const int grid = 148 * 8;
#define RUN_L2(M, name)                                                                          \
  for (int rep = 0; rep < 2; ++rep) {                                                            \
    cudaEventRecord(e0);                                                                         \
    l2_kernel<M><<<grid, 256>>>(table, region_slots, nrows / nregions + 1, nrows, sink);         \
    cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);        \
  }                                                                                              \
  printf("L2  %-44s %7.3f ms  %6.1f G rows/s\n", name, ms, nrows / ms / 1e6);
  RUN_L2(0, "1 RED.f64 / row (region-local, L2-resident)");
  RUN_L2(1, "RED.f64 + RED.u64 same sector");
  RUN_L2(2, "key load + RED.f64 + RED.u64");
  RUN_L2(3, "1 RED.u64 / row");
  RUN_L2(4, "8-byte random load / row");
  RUN_L2(5, "16-byte random load / row");
  const int slots = 8192;
  const size_t smem = (size_t)slots * 20;
  cudaFuncSetAttribute(smem_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(smem_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(smem_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
#define RUN_SM(M, name)                                                                          \
  for (int rep = 0; rep < 2; ++rep) {                                                            \
    cudaEventRecord(e0);                                                                         \
    smem_kernel<M><<<148, 1024, smem>>>(slots, nrows, sink);                                     \
    cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);        \
  }                                                                                              \
  printf("SMEM %-43s %7.3f ms  %6.1f G rows/s\n", name, ms, nrows / ms / 1e6);
  RUN_SM(0, "atomicAdd(double) + atomicAdd(u32), 8K slots");
  RUN_SM(1, "atomicAdd(u32) only");
  RUN_SM(2, "key CAS-probe + atomicAdd(double) + atomicAdd(u32)");
  return 0;
}
