/* CPU oracle (plain C) for the hash -> partition id function and the stable
 * partition it induces.  TEST INFRASTRUCTURE ONLY - see oracle/__init__.py.
 *
 * Restates, for fixed-width key columns:
 *   fugue_dask/_utils.py:146-169   pd.util.hash_pandas_object(df[cols], index=False).mod(num)
 *   pandas/core/util/hashing.py    _hash_ndarray + combine_hash_arrays (pandas 3.0.2,
 *                                  third-party; algorithm quoted in oracle/hash_partition.py)
 * and the "rows of one physical partition are contiguous, input order kept"
 * layout that fugue_dask/_utils.py:124-130 (_postprocess, set_index + repartition)
 * produces per physical partition.
 *
 * Build: make -C oracle   (gcc -O2 -shared -fPIC)  ->  oracle/_build/libfb_oracle.so
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define FBO_NULL_BITS 0x7FF8000000000000ULL

static inline uint64_t fbo_fmix(uint64_t h) {
  h ^= h >> 30; h *= 0xBF58476D1CE4E5B9ULL;
  h ^= h >> 27; h *= 0x94D049BB133111EBULL;
  h ^= h >> 31; return h;
}

static inline uint64_t fbo_bits(const void* p, int width, int64_t i) {
  switch (width) {
    case 1: return ((const uint8_t*)p)[i];
    case 2: return ((const uint16_t*)p)[i];
    case 4: return ((const uint32_t*)p)[i];
    default: return ((const uint64_t*)p)[i];
  }
}

/* out_hash[i] = hash of the key tuple of row i.  valid[k] may be NULL (no nulls)
 * or a byte-per-row mask (0 = NULL). */
void fbo_row_hash(int64_t nrows, int nkeys, const void* const* keys, const int32_t* widths,
                  const uint8_t* const* valid, uint64_t* out_hash) {
  for (int64_t i = 0; i < nrows; ++i) {
    uint64_t out = 0x345678ULL, mult = 1000003ULL;
    for (int k = 0; k < nkeys; ++k) {
      uint64_t b = fbo_bits(keys[k], widths[k], i);
      if (valid && valid[k] && !valid[k][i]) b = FBO_NULL_BITS;
      out ^= fbo_fmix(b);
      out *= mult;
      mult += (uint64_t)(82520 + 2 * (nkeys - k));
    }
    out_hash[i] = out + 97531ULL;
  }
}

void fbo_partition_ids(int64_t nrows, int nkeys, const void* const* keys, const int32_t* widths,
                       const uint8_t* const* valid, uint32_t num, uint32_t* out_pid) {
  uint64_t* h = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(nrows > 0 ? nrows : 1));
  fbo_row_hash(nrows, nkeys, keys, widths, valid, h);
  for (int64_t i = 0; i < nrows; ++i) out_pid[i] = (uint32_t)(h[i] % num);
  free(h);
}

/* Stable counting-sort partition of ncols fixed-width columns.
 * out_offsets has num+1 entries. Returns 0. */
int fbo_partition_cols(int64_t nrows, int ncols, const void* const* cols, const int32_t* col_widths,
                       int nkeys, const int32_t* key_idx, const uint8_t* const* key_valid,
                       uint32_t num, void* const* out_cols, int64_t* out_offsets) {
  const void* keys[16]; int32_t kw[16];
  if (nkeys > 16) return 1;
  for (int k = 0; k < nkeys; ++k) { keys[k] = cols[key_idx[k]]; kw[k] = col_widths[key_idx[k]]; }
  uint32_t* pid = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(nrows > 0 ? nrows : 1));
  int64_t* dst = (int64_t*)malloc(sizeof(int64_t) * (size_t)(nrows > 0 ? nrows : 1));
  fbo_partition_ids(nrows, nkeys, keys, kw, key_valid, num, pid);
  memset(out_offsets, 0, sizeof(int64_t) * ((size_t)num + 1));
  for (int64_t i = 0; i < nrows; ++i) out_offsets[pid[i] + 1]++;
  for (uint32_t p = 0; p < num; ++p) out_offsets[p + 1] += out_offsets[p];
  int64_t* cur = (int64_t*)malloc(sizeof(int64_t) * (size_t)num);
  memcpy(cur, out_offsets, sizeof(int64_t) * (size_t)num);
  for (int64_t i = 0; i < nrows; ++i) dst[i] = cur[pid[i]]++;
  for (int c = 0; c < ncols; ++c) {
    int w = col_widths[c];
    const uint8_t* s = (const uint8_t*)cols[c];
    uint8_t* d = (uint8_t*)out_cols[c];
    if (w == 8) { for (int64_t i = 0; i < nrows; ++i) ((uint64_t*)d)[dst[i]] = ((const uint64_t*)s)[i]; }
    else if (w == 4) { for (int64_t i = 0; i < nrows; ++i) ((uint32_t*)d)[dst[i]] = ((const uint32_t*)s)[i]; }
    else { for (int64_t i = 0; i < nrows; ++i) memcpy(d + dst[i] * w, s + i * w, (size_t)w); }
  }
  free(cur); free(dst); free(pid);
  return 0;
}
