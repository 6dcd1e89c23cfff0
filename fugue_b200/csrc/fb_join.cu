// K7: hash equi-join on sm_100a (build / probe-count / scan / probe-write / gather).
//
// Replaces NativeExecutionEngine.join -> triad PandasUtils.join -> pd.merge
//   fugue/execution/native_execution_engine.py:230-241, schema rule fugue/dataframe/utils.py:152-226.
// NULL keys never match (fugue_test/execution_suite.py:533-543).
//
// Design: open-addressing multimap in HBM, 16-byte slots {key, build_row + 1}; a build row claims
// the first free slot of its probe sequence with one atomicCAS on the row word (duplicates simply
// take further slots, no key comparison while building).  Probing is linear from the same hash;
// pass 1 counts the matches of every probe row, an exclusive scan turns counts into output
// offsets, pass 2 writes (probe_row, build_row) index pairs, and a gather kernel materialises the
// output columns (probe-side indices are monotonic -> coalesced reads; build side is a random
// 8-byte gather).  Output order: probe-row major, deterministic for a given table.
// Algorithmic bytes: 16 + 16 read + 24 written per output row.
#include "fb_common.cuh"

namespace {

struct Slot {
  uint64_t key;
  unsigned long long rowp1;  // 0 = empty
};

inline int64_t l2_batch_bytes() {  // table bytes worked on at a time (B200 L2: 126 MB)
  const char* e = getenv("FB_L2_BATCH_MB");
  const int64_t mb = e != nullptr ? atoll(e) : 64;  // measured 32 / 64 / 96 MB: join 9.53 / 8.91 / 8.74 ms
  return (mb > 0 ? mb : 64) << 20;
}

// clears slots [slot0, slot0 + nslots); status is reset when it is passed
__global__ void fb_join_clear_kernel(Slot* __restrict__ table_all, int64_t slot0, int64_t nslots,
                                     int64_t* __restrict__ status) {
  Slot* __restrict__ table = table_all + slot0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nslots;
       i += (int64_t)gridDim.x * blockDim.x) {
    table[i].key = 0;
    table[i].rowp1 = 0;
  }
  if (status != nullptr && blockIdx.x == 0 && threadIdx.x < 4) status[threadIdx.x] = 0;
}

__global__ void __launch_bounds__(256)
fb_join_build_kernel(const uint64_t* __restrict__ keys, const uint8_t* __restrict__ valid, int64_t n,
                     Slot* __restrict__ table, int64_t capacity, int64_t* __restrict__ status, FbDiv dv,
                     int64_t region_shift, const int64_t* __restrict__ part_off, int p0, int p1) {
  // region_shift >= 0: the table is cut into regions of 1 << region_shift slots, one per hash
  // partition of the (hash-partitioned) inputs, so build and probe sweep it region by region
  // part_off != nullptr: only the rows of hash partitions [p0, p1) (one launch per batch of regions)
  const int64_t mask = region_shift >= 0 ? (((int64_t)1 << region_shift) - 1) : capacity - 1;
  const int64_t row_lo = part_off != nullptr ? part_off[p0] : 0;
  const int64_t row_hi = part_off != nullptr ? part_off[p1] : n;
  for (int64_t i = row_lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < row_hi;
       i += (int64_t)gridDim.x * blockDim.x) {
    if (valid != nullptr && valid[i] == 0) continue;  // NULL keys never match: not inserted
    const uint64_t key = keys[i];
    uint64_t h = fb_fmix64(key);
    int64_t base = 0;
    if (region_shift >= 0) {
      base = (int64_t)fb_fastmod(fb_hash_single_u64(key), dv) << region_shift;
      h >>= 7;
    }
    int64_t s = (int64_t)(h & (uint64_t)mask);
    bool done = false;
    for (int64_t probe = 0; probe <= mask; ++probe) {
      if (table[base + s].rowp1 == 0 &&
          atomicCAS(&table[base + s].rowp1, 0ULL, (unsigned long long)(i + 1)) == 0ULL) {
        table[base + s].key = key;
        done = true;
        break;
      }
      s = (s + 1) & mask;
    }
    if (!done) status[0] = 1;
  }
}

// kWrite == false: counts[i] = matches of probe row i
// kWrite == true : (out_probe, out_build)[offsets[i] + j] = (i, build row of the j-th match);
//                  with `outer`, a probe row without a match emits one pair (i, -1)
template <bool kWrite>
__global__ void __launch_bounds__(256)
fb_join_probe_kernel(const uint64_t* __restrict__ keys, const uint8_t* __restrict__ valid, int64_t n,
                     const Slot* __restrict__ table, int64_t capacity, int outer,
                     int64_t* __restrict__ counts, const int64_t* __restrict__ offsets,
                     int64_t* __restrict__ out_probe, int64_t* __restrict__ out_build, FbDiv dv,
                     int64_t region_shift, int64_t* __restrict__ first) {
  // `first` (optional): the count pass records the build row of the first match (-1: none); the
  // write pass then emits rows with exactly one output pair straight from it, without walking the
  // table again (the common foreign-key -> unique-key join never touches the table twice)
  const int64_t mask = region_shift >= 0 ? (((int64_t)1 << region_shift) - 1) : capacity - 1;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t c = 0;
    int64_t o = kWrite ? offsets[i] : 0;
    int64_t f = -1;
    if (kWrite && first != nullptr) {
      const int64_t cnt = counts[i];
      if (cnt == 0) continue;
      if (cnt == 1) {
        out_probe[o] = i;
        out_build[o] = first[i];
        continue;
      }
    }
    if (valid == nullptr || valid[i] != 0) {
      const uint64_t key = keys[i];
      uint64_t h = fb_fmix64(key);
      int64_t base = 0;
      if (region_shift >= 0) {
        base = (int64_t)fb_fastmod(fb_hash_single_u64(key), dv) << region_shift;
        h >>= 7;
      }
      int64_t s = (int64_t)(h & (uint64_t)mask);
      for (int64_t probe = 0; probe <= mask; ++probe) {
        const unsigned long long r = table[base + s].rowp1;
        if (r == 0) break;
        if (table[base + s].key == key) {
          if (kWrite) {
            out_probe[o + c] = i;
            out_build[o + c] = (int64_t)r - 1;
          } else if (c == 0) {
            f = (int64_t)r - 1;
          }
          ++c;
        }
        s = (s + 1) & mask;
      }
    }
    if (outer && c == 0) {
      if (kWrite) {
        out_probe[o] = i;
        out_build[o] = -1;
      }
      c = 1;
    }
    if (!kWrite) {
      counts[i] = c;
      if (first != nullptr) first[i] = f;
    }
  }
}

// marks build rows that were matched by some probe row (for right/full outer joins)
__global__ void fb_join_mark_kernel(const int64_t* __restrict__ build_idx, int64_t n, uint8_t* __restrict__ matched) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    if (build_idx[i] >= 0) matched[build_idx[i]] = 1;
}

// ---- exclusive scan of int64 (3 kernels: tile sums, scan of sums, tile scan) ------------------
constexpr int kScanBlock = 512;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanBlock * kScanItems;

__device__ __forceinline__ int64_t block_exclusive_scan(int64_t v, int64_t* s_warp, int64_t& total) {
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int64_t x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int64_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
    if (lane >= (unsigned)o) x += y;
  }
  if (lane == 31) s_warp[warp] = x;
  __syncthreads();
  int64_t base = 0, tot = 0;
  for (unsigned w = 0; w < blockDim.x / 32; ++w) {
    int64_t t = s_warp[w];
    if (w < warp) base += t;
    tot += t;
  }
  __syncthreads();
  total = tot;
  return base + x - v;
}

__global__ void __launch_bounds__(kScanBlock)
fb_scan_tile_sums_kernel(const int64_t* __restrict__ in, int64_t n, int64_t* __restrict__ sums) {
  __shared__ int64_t s_warp[kScanBlock / 32];
  const int64_t base = (int64_t)blockIdx.x * kScanTile;
  int64_t v = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    int64_t i = base + (int64_t)k * kScanBlock + threadIdx.x;
    if (i < n) v += in[i];
  }
  int64_t total;
  block_exclusive_scan(v, s_warp, total);
  if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kScanBlock)
fb_scan_sums_kernel(int64_t* __restrict__ sums, int64_t ntiles, int64_t* __restrict__ total_out) {
  __shared__ int64_t s_warp[kScanBlock / 32];
  int64_t carry = 0;
  for (int64_t b0 = 0; b0 < ntiles; b0 += kScanBlock) {
    int64_t i = b0 + threadIdx.x;
    int64_t v = i < ntiles ? sums[i] : 0;
    int64_t total;
    int64_t ex = block_exclusive_scan(v, s_warp, total);
    if (i < ntiles) sums[i] = carry + ex;
    carry += total;
  }
  if (threadIdx.x == 0) *total_out = carry;
}

__global__ void __launch_bounds__(kScanBlock)
fb_scan_tiles_kernel(const int64_t* __restrict__ in, int64_t n, const int64_t* __restrict__ sums,
                     int64_t* __restrict__ out) {
  __shared__ int64_t s_warp[kScanBlock / 32];
  // thread t owns kScanItems consecutive elements
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  int64_t v[kScanItems];
  int64_t sum = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    v[k] = base + k < n ? in[base + k] : 0;
    sum += v[k];
  }
  int64_t total;
  int64_t run = sums[blockIdx.x] + block_exclusive_scan(sum, s_warp, total);
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    if (base + k < n) out[base + k] = run;
    run += v[k];
  }
}

// ---- row gather ----------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
fb_gather_rows_kernel(const void* const* __restrict__ src_cols, void* const* __restrict__ dst_cols,
                      const int32_t* __restrict__ widths, const uint8_t* const* __restrict__ src_valid,
                      uint8_t* const* __restrict__ dst_valid, const int64_t* __restrict__ idx, int64_t n) {
  const int c = blockIdx.y;
  const int w = widths[c];
  const uint8_t* sv = src_valid ? src_valid[c] : nullptr;
  uint8_t* dv = dst_valid ? dst_valid[c] : nullptr;
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n;
       o += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx[o];
    const bool has = r >= 0;
    switch (w) {
      case 8: ((uint64_t*)dst_cols[c])[o] = has ? ((const uint64_t*)src_cols[c])[r] : 0; break;
      case 4: ((uint32_t*)dst_cols[c])[o] = has ? ((const uint32_t*)src_cols[c])[r] : 0; break;
      case 2: ((uint16_t*)dst_cols[c])[o] = has ? ((const uint16_t*)src_cols[c])[r] : 0; break;
      default: ((uint8_t*)dst_cols[c])[o] = has ? ((const uint8_t*)src_cols[c])[r] : 0; break;
    }
    if (dv != nullptr) dv[o] = has ? (sv != nullptr ? sv[r] : (uint8_t)1) : (uint8_t)0;
  }
}

// ---- stream compaction: indices of the non-zero bytes of a mask, in order ---------------------
__global__ void __launch_bounds__(kScanBlock)
fb_mask_tile_counts_kernel(const uint8_t* __restrict__ mask, int64_t n, int64_t* __restrict__ counts) {
  __shared__ int64_t s_warp[kScanBlock / 32];
  const int64_t base = (int64_t)blockIdx.x * kScanTile;
  int64_t v = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    int64_t i = base + (int64_t)k * kScanBlock + threadIdx.x;
    if (i < n && mask[i] != 0) ++v;
  }
  int64_t total;
  block_exclusive_scan(v, s_warp, total);
  if (threadIdx.x == 0) counts[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kScanBlock)
fb_mask_write_kernel(const uint8_t* __restrict__ mask, int64_t n, const int64_t* __restrict__ tile_base,
                     int64_t* __restrict__ out_idx) {
  __shared__ int64_t s_warp[kScanBlock / 32];
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  int64_t cnt = 0;
  uint8_t m[kScanItems];
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    m[k] = base + k < n ? mask[base + k] : 0;
    cnt += m[k] != 0;
  }
  int64_t total;
  int64_t o = tile_base[blockIdx.x] + block_exclusive_scan(cnt, s_warp, total);
#pragma unroll
  for (int k = 0; k < kScanItems; ++k)
    if (m[k] != 0) out_idx[o++] = base + k;
}

inline int64_t region_shift_of(int64_t capacity, uint32_t num_parts) {
  if (num_parts <= 1) return -1;
  int64_t sh = 0;
  while (((int64_t)num_parts << sh) < capacity) ++sh;
  return sh;
}

inline unsigned grid_for(int dev, int64_t n, int per_sm = 8) {
  int64_t b = (n + 255) / 256;
  int64_t m = (int64_t)fb_sm_count(dev) * per_sm;
  if (b > m) b = m;
  if (b < 1) b = 1;
  return (unsigned)b;
}

}  // namespace

extern "C" {

size_t fb_join_table_bytes(int64_t capacity) { return capacity > 0 ? (size_t)capacity * sizeof(Slot) : 0; }

int fb_join_build_u64(int dev, void* stream, int64_t nbuild, const void* keys, const uint8_t* key_valid,
                      int64_t capacity, uint32_t num_parts, void* table, int64_t* d_status,
                      const int64_t* d_part_offsets) {
  FB_CHECK(nbuild >= 0, "nbuild < 0");
  FB_CHECK(capacity >= 2 && (capacity & (capacity - 1)) == 0, "capacity must be a power of two >= 2");
  FB_CHECK(capacity > nbuild, "capacity must exceed the number of build rows");
  FB_CHECK(table != nullptr && d_status != nullptr, "table/status is NULL");
  FB_CHECK(num_parts <= 1 || ((num_parts & (num_parts - 1)) == 0 && (int64_t)num_parts * 2 <= capacity),
           "num_parts must be a power of two <= capacity / 2");
  FbDeviceGuard guard(dev);
  FB_CHECK(guard.ok, "cannot select device %d", dev);
  cudaStream_t st = (cudaStream_t)stream;
  const FbDiv dv = fb_make_div(num_parts > 1 ? num_parts : 1);
  const int64_t rs = region_shift_of(capacity, num_parts);
  if (num_parts > 1 && d_part_offsets != nullptr && nbuild > 0) {
    // clear + fill a few regions at a time, so that they are still in L2 when the inserts arrive
    // (a table cleared as a whole is back in HBM by then: 5.5 ms for 62.5 M rows, one random DRAM
    // sector read + write-back per insert)
    const int64_t region_bytes = ((int64_t)1 << rs) * (int64_t)sizeof(Slot);
    int64_t per = l2_batch_bytes() / region_bytes;
    if (per < 1) per = 1;
    FB_CUDA(cudaMemsetAsync(d_status, 0, 4 * sizeof(int64_t), st));
    for (int64_t p0 = 0; p0 < (int64_t)num_parts; p0 += per) {
      const int64_t p1 = p0 + per < (int64_t)num_parts ? p0 + per : (int64_t)num_parts;
      const int64_t nslots = (p1 - p0) << rs;
      fb_join_clear_kernel<<<grid_for(dev, nslots / 4 + 1), 256, 0, st>>>((Slot*)table, p0 << rs, nslots, nullptr);
      const int64_t est = nbuild / num_parts * (p1 - p0) * 5 / 4 + 256;
      fb_join_build_kernel<<<grid_for(dev, est), 256, 0, st>>>((const uint64_t*)keys, key_valid, nbuild, (Slot*)table,
                                                              capacity, d_status, dv, rs, d_part_offsets, (int)p0,
                                                              (int)p1);
    }
    FB_CUDA(cudaGetLastError());
    return 0;
  }
  fb_join_clear_kernel<<<grid_for(dev, capacity), 256, 0, st>>>((Slot*)table, 0, capacity, d_status);
  FB_CUDA(cudaGetLastError());
  if (nbuild > 0) {
    fb_join_build_kernel<<<grid_for(dev, nbuild), 256, 0, st>>>((const uint64_t*)keys, key_valid, nbuild,
                                                                (Slot*)table, capacity, d_status, dv, rs, nullptr, 0, 0);
    FB_CUDA(cudaGetLastError());
  }
  return 0;
}

int fb_join_probe_count_u64(int dev, void* stream, int64_t nprobe, const void* keys,
                            const uint8_t* key_valid, int64_t capacity, uint32_t num_parts,
                            const void* table, int outer, int64_t* out_counts, int64_t* out_first) {
  FB_CHECK(nprobe >= 0, "nprobe < 0");
  if (nprobe == 0) return 0;
  FbDeviceGuard guard(dev);
  FB_CHECK(guard.ok, "cannot select device %d", dev);
  fb_join_probe_kernel<false><<<grid_for(dev, nprobe), 256, 0, (cudaStream_t)stream>>>(
      (const uint64_t*)keys, key_valid, nprobe, (const Slot*)table, capacity, outer, out_counts, nullptr,
      nullptr, nullptr, fb_make_div(num_parts > 1 ? num_parts : 1), region_shift_of(capacity, num_parts),
      out_first);
  FB_CUDA(cudaGetLastError());
  return 0;
}

int fb_join_probe_write_u64(int dev, void* stream, int64_t nprobe, const void* keys,
                            const uint8_t* key_valid, int64_t capacity, uint32_t num_parts,
                            const void* table, int outer, const int64_t* offsets, int64_t* out_probe_idx,
                            int64_t* out_build_idx, const int64_t* counts, const int64_t* first) {
  FB_CHECK(nprobe >= 0, "nprobe < 0");
  if (nprobe == 0) return 0;
  FbDeviceGuard guard(dev);
  FB_CHECK(guard.ok, "cannot select device %d", dev);
  fb_join_probe_kernel<true><<<grid_for(dev, nprobe), 256, 0, (cudaStream_t)stream>>>(
      (const uint64_t*)keys, key_valid, nprobe, (const Slot*)table, capacity, outer,
      (counts != nullptr && first != nullptr) ? (int64_t*)counts : nullptr, offsets, out_probe_idx, out_build_idx,
      fb_make_div(num_parts > 1 ? num_parts : 1), region_shift_of(capacity, num_parts),
      (counts != nullptr && first != nullptr) ? (int64_t*)first : nullptr);
  FB_CUDA(cudaGetLastError());
  return 0;
}

int fb_join_mark_matched(int dev, void* stream, const int64_t* build_idx, int64_t n, uint8_t* matched) {
  if (n <= 0) return 0;
  FbDeviceGuard guard(dev);
  FB_CHECK(guard.ok, "cannot select device %d", dev);
  fb_join_mark_kernel<<<grid_for(dev, n), 256, 0, (cudaStream_t)stream>>>(build_idx, n, matched);
  FB_CUDA(cudaGetLastError());
  return 0;
}

size_t fb_exclusive_scan_scratch_bytes(int64_t n) {
  int64_t ntiles = (n + kScanTile - 1) / kScanTile;
  return (size_t)(ntiles + 1) * sizeof(int64_t);
}

int fb_exclusive_scan_i64(int dev, void* stream, int64_t n, const int64_t* in, int64_t* out,
                          int64_t* out_total, void* scratch, size_t scratch_bytes) {
  FB_CHECK(n >= 0, "n < 0");
  FB_CHECK(out_total != nullptr, "out_total is NULL");
  FbDeviceGuard guard(dev);
  FB_CHECK(guard.ok, "cannot select device %d", dev);
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) {
    FB_CUDA(cudaMemsetAsync(out_total, 0, sizeof(int64_t), st));
    return 0;
  }
  FB_CHECK(scratch != nullptr && scratch_bytes >= fb_exclusive_scan_scratch_bytes(n), "scan scratch too small");
  const int64_t ntiles = (n + kScanTile - 1) / kScanTile;
  int64_t* sums = (int64_t*)scratch;
  fb_scan_tile_sums_kernel<<<(unsigned)ntiles, kScanBlock, 0, st>>>(in, n, sums);
  FB_CUDA(cudaGetLastError());
  fb_scan_sums_kernel<<<1, kScanBlock, 0, st>>>(sums, ntiles, out_total);
  FB_CUDA(cudaGetLastError());
  fb_scan_tiles_kernel<<<(unsigned)ntiles, kScanBlock, 0, st>>>(in, n, sums, out);
  FB_CUDA(cudaGetLastError());
  return 0;
}

size_t fb_compact_scratch_bytes(int64_t n) { return fb_exclusive_scan_scratch_bytes(n); }

int fb_compact_indices(int dev, void* stream, const uint8_t* mask, int64_t n, int64_t* out_idx,
                       int64_t* d_count, void* scratch, size_t scratch_bytes) {
  FB_CHECK(n >= 0 && d_count != nullptr, "bad arguments");
  FbDeviceGuard guard(dev);
  FB_CHECK(guard.ok, "cannot select device %d", dev);
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) {
    FB_CUDA(cudaMemsetAsync(d_count, 0, sizeof(int64_t), st));
    return 0;
  }
  FB_CHECK(scratch != nullptr && scratch_bytes >= fb_compact_scratch_bytes(n), "compaction scratch too small");
  const int64_t ntiles = (n + kScanTile - 1) / kScanTile;
  int64_t* counts = (int64_t*)scratch;
  fb_mask_tile_counts_kernel<<<(unsigned)ntiles, kScanBlock, 0, st>>>(mask, n, counts);
  FB_CUDA(cudaGetLastError());
  fb_scan_sums_kernel<<<1, kScanBlock, 0, st>>>(counts, ntiles, d_count);
  FB_CUDA(cudaGetLastError());
  fb_mask_write_kernel<<<(unsigned)ntiles, kScanBlock, 0, st>>>(mask, n, counts, out_idx);
  FB_CUDA(cudaGetLastError());
  return 0;
}

int fb_gather_rows(int dev, void* stream, int ncols, const void* const* d_src_cols, void* const* d_dst_cols,
                   const int32_t* d_widths, const uint8_t* const* d_src_valid, uint8_t* const* d_dst_valid,
                   const int64_t* idx, int64_t n) {
  FB_CHECK(ncols >= 0 && n >= 0, "negative count");
  if (ncols == 0 || n == 0) return 0;
  FbDeviceGuard guard(dev);
  FB_CHECK(guard.ok, "cannot select device %d", dev);
  dim3 grid(grid_for(dev, n, 4), (unsigned)ncols);
  fb_gather_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_src_cols, d_dst_cols, d_widths, d_src_valid,
                                                               d_dst_valid, idx, n);
  FB_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
