// Segment copy: out[dst_off[s] .. +len[s]) = table[src_tab[s]][src_off[s] .. +len[s]) for every column.
//
// Multi-GPU exchange epilogue (SURVEY.md 8e steps 3+4 in one kernel): the source tables are the
// hash-partitioned column buffers of ALL ranks, mapped into this process through symmetric memory
// (NVLink peer pointers), so the kernel PULLS every (source rank, partition) run straight from the
// peer's HBM over NVLink 5 into its final place in the owned, partition-contiguous output - no
// NCCL all-to-all, no staging pass.  With one source table it is a plain local segment copy.
// The reference has no counterpart (its shuffles are Dask/Spark/Ray's, fugue_dask/_utils.py:124-130).
// Pure byte movement: 2 x width bytes per row per column; peer loads need many bytes in flight
// (NVLink latency ~3x HBM), hence 8 independent 8-byte loads per thread and piece-parallel CTAs.
#include "fb_common.cuh"

namespace {

constexpr int kPieceRows = 8192;  // rows one CTA handles per piece step

template <typename T, int kCopyBlock, int kCopyUnroll>
__device__ __forceinline__ void copy_run(const T* __restrict__ src, T* __restrict__ dst, int64_t n,
                                         int piece, int npieces) {
  for (int64_t base = (int64_t)piece * kPieceRows; base < n; base += (int64_t)npieces * kPieceRows) {
    const int64_t end = base + kPieceRows < n ? base + kPieceRows : n;
    int64_t i = base + threadIdx.x;
    for (; i + (kCopyUnroll - 1) * kCopyBlock < end; i += kCopyUnroll * kCopyBlock) {
      T v[kCopyUnroll];
#pragma unroll
      for (int k = 0; k < kCopyUnroll; ++k) v[k] = src[i + k * kCopyBlock];
#pragma unroll
      for (int k = 0; k < kCopyUnroll; ++k) dst[i + k * kCopyBlock] = v[k];
    }
    for (; i < end; i += kCopyBlock) dst[i] = src[i];
  }
}

// 8-byte elements with 16-byte loads: peer (NVLink) reads are measured ~35 % faster with 16-byte
// requests per lane than with 8-byte ones.  The source run is aligned to 16 bytes by peeling one
// element; the destination is local HBM and is written with 8-byte stores (its alignment is free).
template <int kCopyBlock, int kCopyUnroll>
__device__ __forceinline__ void copy_run_u64_v2(const uint64_t* __restrict__ src, uint64_t* __restrict__ dst,
                                                int64_t n, int piece, int npieces) {
  int64_t head = ((uintptr_t)src & 8) ? 1 : 0;
  if (head > n) head = n;
  if (piece == 0 && head == 1 && threadIdx.x == 0) dst[0] = src[0];
  const ulonglong2* __restrict__ s2 = (const ulonglong2*)(src + head);
  uint64_t* __restrict__ d = dst + head;
  const int64_t npairs = (n - head) >> 1;
  constexpr int64_t kPiecePairs = kPieceRows / 2;
  for (int64_t base = (int64_t)piece * kPiecePairs; base < npairs; base += (int64_t)npieces * kPiecePairs) {
    const int64_t end = base + kPiecePairs < npairs ? base + kPiecePairs : npairs;
    int64_t i = base + threadIdx.x;
    for (; i + (kCopyUnroll - 1) * kCopyBlock < end; i += kCopyUnroll * kCopyBlock) {
      ulonglong2 v[kCopyUnroll];
#pragma unroll
      for (int k = 0; k < kCopyUnroll; ++k) v[k] = s2[i + k * kCopyBlock];
#pragma unroll
      for (int k = 0; k < kCopyUnroll; ++k) {
        d[2 * (i + k * kCopyBlock)] = v[k].x;
        d[2 * (i + k * kCopyBlock) + 1] = v[k].y;
      }
    }
    for (; i < end; i += kCopyBlock) {
      const ulonglong2 v = s2[i];
      d[2 * i] = v.x;
      d[2 * i + 1] = v.y;
    }
  }
  if (piece == 0 && ((n - head) & 1) && threadIdx.x == 0) dst[n - 1] = src[n - 1];
}

template <int kCopyBlock, int kCopyUnroll>
__global__ void __launch_bounds__(kCopyBlock)
fb_copy_segments_kernel(const void* const* __restrict__ src_cols, void* const* __restrict__ dst_cols,
                        const int32_t* __restrict__ widths, int ncols, const int32_t* __restrict__ src_tab,
                        const int64_t* __restrict__ src_off, const int64_t* __restrict__ dst_off,
                        const int64_t* __restrict__ len, int nseg) {
  const int c = blockIdx.y;
  const int w = widths[c];
  uint8_t* d = (uint8_t*)dst_cols[c];
  const int piece = blockIdx.z, npieces = gridDim.z;
  for (int sgi = blockIdx.x; sgi < nseg; sgi += gridDim.x) {
    const int64_t n = len[sgi];
    if (n <= 0) continue;
    const int tab = src_tab != nullptr ? src_tab[sgi] : 0;
    const uint8_t* s = (const uint8_t*)src_cols[(size_t)tab * ncols + c];
    const int64_t so = src_off[sgi], dof = dst_off[sgi];
    switch (w) {
      case 8: copy_run_u64_v2<kCopyBlock, kCopyUnroll>((const uint64_t*)s + so, (uint64_t*)d + dof, n, piece, npieces); break;
      case 4: copy_run<uint32_t, kCopyBlock, kCopyUnroll>((const uint32_t*)s + so, (uint32_t*)d + dof, n, piece, npieces); break;
      case 2: copy_run<uint16_t, kCopyBlock, kCopyUnroll>((const uint16_t*)s + so, (uint16_t*)d + dof, n, piece, npieces); break;
      default: copy_run<uint8_t, kCopyBlock, kCopyUnroll>(s + so, d + dof, n, piece, npieces); break;
    }
  }
}

}  // namespace

extern "C" int fb_copy_segments(int dev, void* stream, int ncols, const void* const* d_src_cols,
                                void* const* d_dst_cols, const int32_t* d_widths, int nseg,
                                const int32_t* d_src_table, const int64_t* d_src_off,
                                const int64_t* d_dst_off, const int64_t* d_len, int64_t max_len) {
  FB_CHECK(ncols >= 0 && nseg >= 0, "negative count");
  if (ncols == 0 || nseg == 0) return 0;
  FbDeviceGuard guard(dev);
  FB_CHECK(guard.ok, "cannot select device %d", dev);
  const int sms = fb_sm_count(dev);
  int gx = nseg < 65535 ? nseg : 65535;
  // enough pieces to give every SM several CTAs even with few, long segments
  int64_t pieces = max_len > 0 ? (max_len + kPieceRows - 1) / kPieceRows : 1;
  const int64_t want = (int64_t)sms * 16 / ((int64_t)gx * ncols) + 1;
  if (pieces > want) pieces = want;
  if (pieces > 64) pieces = 64;
  if (pieces < 1) pieces = 1;
  dim3 grid((unsigned)gx, (unsigned)ncols, (unsigned)pieces);
  // 256 threads x 8 independent 16-byte loads per thread; other shapes measured within 3 %
  fb_copy_segments_kernel<256, 8><<<grid, 256, 0, (cudaStream_t)stream>>>(
      d_src_cols, d_dst_cols, d_widths, ncols, d_src_table, d_src_off, d_dst_off, d_len, nseg);
  FB_CUDA(cudaGetLastError());
  return 0;
}


// ---------------------------------------------------------------------------
// TMA pull: the exchange as a small persistent kernel.  Each CTA owns a ring of 16 KB shared-memory
// stages; ONE thread keeps `nstages` bulk loads (cp.async.bulk global -> shared, SASS UBLKCP) in
// flight against the peers' HBM over NVLink - ~190 KB per SM without a register staging, so a
// handful of SMs cover the NVLink latency x bandwidth product - and four consumer warps drain the
// stages into local HBM with plain 8-byte stores (the destination run may be only 8-byte aligned;
// bulk stores need 16).  Runs are cut into chunks that are dealt round-robin to the CTAs, so every
// CTA reads from every peer.  src / dst / bytes must be multiples of 8.
// ---------------------------------------------------------------------------
namespace {

constexpr int kPullChunk = 16384;       // bytes per bulk load
constexpr int kPullMaxRuns = 64;
constexpr int kPullConsumers = 128;     // 4 warps
constexpr int kPullThreads = kPullConsumers + 32;

struct PullRuns {
  const uint8_t* src[kPullMaxRuns];   // 16-byte aligned body start
  uint8_t* dst[kPullMaxRuns];
  uint64_t body[kPullMaxRuns];        // bytes of the 16-byte-multiple body
  uint32_t first_chunk[kPullMaxRuns + 1];
  // fragments outside the body (at most one 8-byte head and one 8-byte tail per run)
  const uint64_t* frag_src[2 * kPullMaxRuns];
  uint64_t* frag_dst[2 * kPullMaxRuns];
  int32_t nruns, nfrags;
};

__device__ __forceinline__ uint32_t pl_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void pl_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nPL_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra PL_DONE;\nbra PL_WAIT;\nPL_DONE:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}

__global__ void __launch_bounds__(kPullThreads, 1)
fb_pull_tma_kernel(const __grid_constant__ PullRuns runs, int nstages) {
  extern __shared__ __align__(128) uint8_t pl_ring[];
  __shared__ __align__(8) uint64_t bars[32];
  const uint32_t full = pl_smem(bars), empty = pl_smem(bars + 16);
  if (threadIdx.x == 0) {
    for (int s = 0; s < nstages; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(full + 8 * s), "r"(1) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(empty + 8 * s), "r"(kPullConsumers / 32) : "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const uint32_t total = runs.first_chunk[runs.nruns];
  if (threadIdx.x >= kPullConsumers) {
    if (threadIdx.x != kPullConsumers) return;
    // ---- producer: one thread
    uint32_t s = 0, ph = 0;
    int r = 0;
    for (uint32_t g = blockIdx.x; g < total; g += gridDim.x) {
      while (g >= runs.first_chunk[r + 1]) ++r;
      const uint64_t off = (uint64_t)(g - runs.first_chunk[r]) * kPullChunk;
      const uint64_t left = runs.body[r] - off;
      const uint32_t nb = left < kPullChunk ? (uint32_t)left : (uint32_t)kPullChunk;
      pl_wait(empty + 8 * s, ph ^ 1);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full + 8 * s), "r"(nb) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(pl_smem(pl_ring) + s * kPullChunk), "l"(runs.src[r] + off), "r"(nb), "r"(full + 8 * s)
                   : "memory");
      if (++s == (uint32_t)nstages) { s = 0; ph ^= 1; }
    }
    return;
  }
  // ---- consumers
  if (blockIdx.x == 0)
    for (int f = threadIdx.x; f < runs.nfrags; f += kPullConsumers) *runs.frag_dst[f] = *runs.frag_src[f];
  uint32_t s = 0, ph = 0;
  int r = 0;
  const unsigned lane = threadIdx.x & 31;
  for (uint32_t g = blockIdx.x; g < total; g += gridDim.x) {
    while (g >= runs.first_chunk[r + 1]) ++r;
    const uint64_t off = (uint64_t)(g - runs.first_chunk[r]) * kPullChunk;
    const uint64_t left = runs.body[r] - off;
    const uint32_t n8 = (left < kPullChunk ? (uint32_t)left : (uint32_t)kPullChunk) >> 3;
    pl_wait(full + 8 * s, ph);
    const uint64_t* __restrict__ st = (const uint64_t*)(pl_ring + (size_t)s * kPullChunk);
    uint64_t* __restrict__ d = (uint64_t*)(runs.dst[r] + off);
#pragma unroll 4
    for (uint32_t i = threadIdx.x; i < n8; i += kPullConsumers) d[i] = st[i];
    __syncwarp();
    if (lane == 0)
      asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(empty + 8 * s) : "memory");
    if (++s == (uint32_t)nstages) { s = 0; ph ^= 1; }
  }
}

}  // namespace

extern "C" int fb_pull_runs_tma(int dev, void* stream, int nruns, const void* const* src, void* const* dst,
                                const size_t* bytes, int max_ctas) {
  FB_CHECK(nruns >= 0 && nruns <= kPullMaxRuns, "nruns=%d out of range [0,%d]", nruns, kPullMaxRuns);
  if (nruns == 0) return 0;
  FB_CHECK(src != nullptr && dst != nullptr && bytes != nullptr, "NULL argument");
  FbDeviceGuard guard(dev);
  FB_CHECK(guard.ok, "cannot select device %d", dev);
  PullRuns pr;
  memset(&pr, 0, sizeof(pr));
  uint32_t chunks = 0;
  for (int i = 0; i < nruns; ++i) {
    uintptr_t s = (uintptr_t)src[i], d = (uintptr_t)dst[i];
    uint64_t n = bytes[i];
    FB_CHECK(s % 8 == 0 && d % 8 == 0 && n % 8 == 0, "run %d: pointers and size must be multiples of 8", i);
    if (n >= 8 && (s & 8)) {  // head element brings the source to 16-byte alignment
      pr.frag_src[pr.nfrags] = (const uint64_t*)s;
      pr.frag_dst[pr.nfrags++] = (uint64_t*)d;
      s += 8; d += 8; n -= 8;
    }
    if (n & 8) {  // bulk sizes are multiples of 16
      pr.frag_src[pr.nfrags] = (const uint64_t*)(s + n - 8);
      pr.frag_dst[pr.nfrags++] = (uint64_t*)(d + n - 8);
      n -= 8;
    }
    pr.src[i] = (const uint8_t*)s;
    pr.dst[i] = (uint8_t*)d;
    pr.body[i] = n;
    pr.first_chunk[i] = chunks;
    const uint64_t c = (n + kPullChunk - 1) / kPullChunk;
    FB_CHECK(chunks + c < ((uint64_t)1 << 32), "too many chunks");
    chunks += (uint32_t)c;
  }
  pr.first_chunk[nruns] = chunks;
  pr.nruns = nruns;
  int smem_max = 0;
  FB_CUDA(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  int nstages = (smem_max - 1024) / kPullChunk;
  if (nstages > 14) nstages = 14;
  FB_CHECK(nstages >= 2, "not enough shared memory");
  const size_t smem = (size_t)nstages * kPullChunk;
  static bool optin_done = false;  // same value every time; a race only repeats the call
  if (!optin_done) {
    FB_CUDA(cudaFuncSetAttribute(fb_pull_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max - 1024));
    optin_done = true;
  }
  int grid = max_ctas > 0 ? max_ctas : 16;
  if ((uint32_t)grid > chunks) grid = chunks > 0 ? (int)chunks : 1;
  fb_pull_tma_kernel<<<grid, kPullThreads, smem, (cudaStream_t)stream>>>(pr, nstages);
  FB_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------
// The exchange on the copy engines (see include/fugue_b200.h): a few large runs, one
// cudaMemcpyAsync each (peer memory mapped through symmetric memory is a plain device pointer
// here; the driver routes the copy over NVLink).
// ---------------------------------------------------------------------------
extern "C" int fb_copy_runs_dma(int dev, void* stream, int64_t nruns, const void* const* src, void* const* dst,
                                const size_t* bytes) {
  FB_CHECK(nruns >= 0, "nruns < 0");
  if (nruns == 0) return 0;
  FB_CHECK(src != nullptr && dst != nullptr && bytes != nullptr, "NULL argument");
  FbDeviceGuard guard(dev);
  FB_CHECK(guard.ok, "cannot select device %d", dev);
  cudaStream_t st = (cudaStream_t)stream;
  for (int64_t i = 0; i < nruns; ++i)
    if (bytes[i] != 0) FB_CUDA(cudaMemcpyAsync(dst[i], src[i], bytes[i], cudaMemcpyDeviceToDevice, st));
  return 0;
}

extern "C" int fb_copy_runs_dma_streams(int dev, int64_t nruns, const void* const* src, void* const* dst,
                                        const size_t* bytes, void* const* streams, int prefer_overlap) {
  FB_CHECK(nruns >= 0, "nruns < 0");
  if (nruns == 0) return 0;
  FB_CHECK(src != nullptr && dst != nullptr && bytes != nullptr && streams != nullptr, "NULL argument");
  FbDeviceGuard guard(dev);
  FB_CHECK(guard.ok, "cannot select device %d", dev);
  for (int64_t i = 0; i < nruns; ++i) {
    if (bytes[i] == 0) continue;
#if CUDART_VERSION >= 12080
    if (prefer_overlap) {
      // cudaMemcpyFlagPreferOverlapWithCompute: "try and overlap the copy with compute work on the SMs" - without
      // it the copies made no progress next to the persistent scatter kernel (profiles/r2_exchange_notes.md)
      cudaMemcpyAttributes attr;
      memset(&attr, 0, sizeof(attr));
      attr.srcAccessOrder = cudaMemcpySrcAccessOrderStream;
      attr.flags = cudaMemcpyFlagPreferOverlapWithCompute;
      void* d = dst[i];
      void* s_ = (void*)src[i];
      size_t n = bytes[i], idx = 0, fail = 0;
      cudaError_t e = cudaMemcpyBatchAsync(&d, &s_, &n, 1, &attr, &idx, 1, &fail, (cudaStream_t)streams[i]);
      if (e == cudaSuccess) continue;
      (void)cudaGetLastError();  // not supported for these pointers: plain copy below
    }
#endif
    FB_CUDA(cudaMemcpyAsync(dst[i], src[i], bytes[i], cudaMemcpyDeviceToDevice, (cudaStream_t)streams[i]));
  }
  return 0;
}
