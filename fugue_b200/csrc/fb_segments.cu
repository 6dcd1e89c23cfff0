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
