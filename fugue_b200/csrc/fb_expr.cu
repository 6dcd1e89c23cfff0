// Column-expression evaluator for sm_100a: one pass over HBM per SELECT list.
//
// ExecutionEngine.select / filter / assign (fugue/execution/execution_engine.py:736-887) evaluate
// expression trees (fugue/column/expressions.py); the reference turns them into SQL text and hands
// them to qpd/pandas, which materialises one temporary column per operator.  Here the host compiles
// the whole SELECT list (or WHERE predicate) into one short accumulator-machine program:
//
//   * a thread owns kExprItems rows; the accumulator (their current values + a validity bit each)
//     sits in hardware registers for the whole program;
//   * the second operand of an instruction is fetched straight from its source: a column in HBM
//     (coalesced, converted from its storage type on the fly), an immediate, or a temporary in
//     shared memory - temporaries are only needed when both sides of an operator are compound, so
//     typical expressions never touch shared memory (a first version that kept every intermediate
//     in shared-memory vector registers ran at 1.0 TB/s, bound by shared-memory wavefronts);
//   * the interpreter dispatches once per instruction per thread and then runs its rows unrolled,
//     so the switch is amortised; every output is written by its own FB_X_OUT instruction;
//   * SQL NULL semantics: arithmetic and comparisons propagate NULL (validity bits are ANDed for
//     all rows of the thread at once), AND / OR are Kleene three-valued, IS NULL / IS NOT NULL /
//     COALESCE read the validity bits.
#include <mutex>

#include "fb_common.cuh"

namespace {

constexpr int kExprThreads = 256;
constexpr int kExprItems = 8;
constexpr int kExprTile = kExprThreads * kExprItems;
constexpr unsigned kAllValid = (1u << kExprItems) - 1;

struct ExprProgram {
  const void* col_ptr[FB_EXPR_MAX_COLS];
  const uint8_t* col_valid[FB_EXPR_MAX_COLS];
  int32_t col_type[FB_EXPR_MAX_COLS];
  void* out_ptr[FB_EXPR_MAX_OUTS];
  uint8_t* out_valid[FB_EXPR_MAX_OUTS];
  int32_t out_type[FB_EXPR_MAX_OUTS];
  int32_t nins;
  fb_expr_ins ins[FB_EXPR_MAX_INS];
};
static_assert(sizeof(ExprProgram) <= 4000, "program must fit the kernel parameter space");

__device__ __forceinline__ double as_f(uint64_t b) { return __longlong_as_double((long long)b); }
__device__ __forceinline__ uint64_t f_bits(double d) { return (uint64_t)__double_as_longlong(d); }

template <typename T, bool kFloat, bool kFull>
__device__ __forceinline__ void load_col(const void* p, int64_t row0, int tid, int nk, uint64_t (&v)[kExprItems]) {
  const T* __restrict__ q = (const T*)p + row0;
#pragma unroll
  for (int k = 0; k < kExprItems; ++k) {
    const int i = k * kExprThreads + tid;
    if (kFull || i < nk) {
      if (kFloat) v[k] = f_bits((double)q[i]);
      else v[k] = (uint64_t)(int64_t)q[i];
    }
  }
}

template <typename T, bool kFloat, bool kFull>
__device__ __forceinline__ void store_col(void* p, int64_t row0, int tid, int nk, const uint64_t (&v)[kExprItems],
                                          unsigned valid) {
  T* __restrict__ q = (T*)p + row0;
#pragma unroll
  for (int k = 0; k < kExprItems; ++k) {
    const int i = k * kExprThreads + tid;
    if (kFull || i < nk) {
      const uint64_t bits = (valid >> k) & 1u ? v[k] : 0ull;  // NULL rows store 0
      if (kFloat) q[i] = (T)as_f(bits);
      else q[i] = (T)(int64_t)bits;
    }
  }
}

// one tile of kExprTile rows (kFull: no bounds checks; only the last tile of a table is partial)
template <bool kFull>
__device__ __forceinline__ void run_tile(const ExprProgram& P, int64_t row0, int nk, int tid, uint64_t* tmp_v,
                                         uint8_t* tmp_m) {
  {
    uint64_t acc[kExprItems];
    unsigned accv = kAllValid;
#pragma unroll
    for (int k = 0; k < kExprItems; ++k) acc[k] = 0;
    for (int pc = 0; pc < P.nins; ++pc) {
      const fb_expr_ins in = P.ins[pc];
      // ---- operand B
      uint64_t b[kExprItems];
      unsigned bv = kAllValid;
      if (in.kind == FB_XK_IMM) {
#pragma unroll
        for (int k = 0; k < kExprItems; ++k) b[k] = (uint64_t)in.imm;
      } else if (in.kind == FB_XK_COL) {
        const void* p = P.col_ptr[in.b];
        switch (P.col_type[in.b]) {
          case FB_T_I8: load_col<int8_t, false, kFull>(p, row0, tid, nk, b); break;
          case FB_T_I16: load_col<int16_t, false, kFull>(p, row0, tid, nk, b); break;
          case FB_T_I32: load_col<int32_t, false, kFull>(p, row0, tid, nk, b); break;
          case FB_T_I64: load_col<int64_t, false, kFull>(p, row0, tid, nk, b); break;
          case FB_T_U8: load_col<uint8_t, false, kFull>(p, row0, tid, nk, b); break;
          case FB_T_F32: load_col<float, true, kFull>(p, row0, tid, nk, b); break;
          default: load_col<int64_t, false, kFull>(p, row0, tid, nk, b); break;  // FB_T_F64: raw bits
        }
        const uint8_t* m = P.col_valid[in.b];
        if (m != nullptr) {
          bv = 0;
#pragma unroll
          for (int k = 0; k < kExprItems; ++k) {
            const int i = k * kExprThreads + tid;
            if ((kFull || i < nk) && m[row0 + i] != 0) bv |= 1u << k;
          }
        }
      } else if (in.kind == FB_XK_REG) {
        const uint64_t* r = tmp_v + (size_t)in.b * kExprTile;
#pragma unroll
        for (int k = 0; k < kExprItems; ++k) b[k] = r[k * kExprThreads + tid];
        bv = tmp_m[in.b * kExprThreads + tid];
      } else if (in.kind == FB_XK_NULL) {
        bv = 0;
      }
      if (in.kind == FB_XK_NONE || in.kind == FB_XK_NULL || (!kFull && in.kind == FB_XK_COL)) {
#pragma unroll
        for (int k = 0; k < kExprItems; ++k)  // defined values for lanes that were not loaded
          if (in.kind != FB_XK_COL || k * kExprThreads + tid >= nk) b[k] = 0;
      }
      if (in.flags & FB_XF_B_I2F) {
#pragma unroll
        for (int k = 0; k < kExprItems; ++k) b[k] = f_bits((double)(int64_t)b[k]);
      }
#define FB_ROWS(...) _Pragma("unroll") for (int k = 0; k < kExprItems; ++k) { __VA_ARGS__ }
#define FB_BIN(EXPR) FB_ROWS(const uint64_t x = acc[k], y = b[k]; acc[k] = (EXPR);) accv &= bv;
#define FB_UN(EXPR) FB_ROWS(const uint64_t x = acc[k]; acc[k] = (EXPR);)
      switch (in.op) {
        case FB_X_MOV: FB_ROWS(acc[k] = b[k];) accv = bv; break;
        case FB_X_ST: {
          uint64_t* r = tmp_v + (size_t)in.b * kExprTile;
          FB_ROWS(r[k * kExprThreads + tid] = acc[k];)
          tmp_m[in.b * kExprThreads + tid] = (uint8_t)accv;
          break;
        }
        case FB_X_OUT: {
          void* p = P.out_ptr[in.b];
          switch (P.out_type[in.b]) {
            case FB_T_I8: store_col<int8_t, false, kFull>(p, row0, tid, nk, acc, accv); break;
            case FB_T_I16: store_col<int16_t, false, kFull>(p, row0, tid, nk, acc, accv); break;
            case FB_T_I32: store_col<int32_t, false, kFull>(p, row0, tid, nk, acc, accv); break;
            case FB_T_I64: store_col<int64_t, false, kFull>(p, row0, tid, nk, acc, accv); break;
            case FB_T_U8: store_col<uint8_t, false, kFull>(p, row0, tid, nk, acc, accv); break;
            case FB_T_F32: store_col<float, true, kFull>(p, row0, tid, nk, acc, accv); break;
            default: store_col<int64_t, false, kFull>(p, row0, tid, nk, acc, accv); break;  // FB_T_F64
          }
          uint8_t* m = P.out_valid[in.b];
          if (m != nullptr) {
            FB_ROWS(const int i = k * kExprThreads + tid;
                    if (kFull || i < nk) m[row0 + i] = (uint8_t)((accv >> k) & 1u);)
          }
          break;
        }
        case FB_X_I2F: FB_UN(f_bits((double)(int64_t)x)) break;
        case FB_X_F2I: FB_UN((uint64_t)(int64_t)as_f(x)) break;
        case FB_X_NEG_I: FB_UN(0 - x) break;
        case FB_X_NEG_F: FB_UN(x ^ 0x8000000000000000ull) break;
        case FB_X_NOT: FB_UN((uint64_t)(x == 0)) break;
        case FB_X_IS_NULL: FB_ROWS(acc[k] = (uint64_t)(((accv >> k) & 1u) == 0);) accv = kAllValid; break;
        case FB_X_NOT_NULL: FB_ROWS(acc[k] = (uint64_t)((accv >> k) & 1u);) accv = kAllValid; break;
        case FB_X_TOBOOL_I: FB_UN((uint64_t)(x != 0)) break;
        case FB_X_TOBOOL_F: FB_UN((uint64_t)(as_f(x) != 0.0)) break;
        case FB_X_ADD_I: FB_BIN(x + y) break;
        case FB_X_SUB_I: FB_BIN(x - y) break;
        case FB_X_RSUB_I: FB_BIN(y - x) break;
        case FB_X_MUL_I: FB_BIN(x * y) break;
        case FB_X_ADD_F: FB_BIN(f_bits(as_f(x) + as_f(y))) break;
        case FB_X_SUB_F: FB_BIN(f_bits(as_f(x) - as_f(y))) break;
        case FB_X_RSUB_F: FB_BIN(f_bits(as_f(y) - as_f(x))) break;
        case FB_X_MUL_F: FB_BIN(f_bits(as_f(x) * as_f(y))) break;
        case FB_X_DIV_F: FB_BIN(f_bits(as_f(x) / as_f(y))) break;
        case FB_X_RDIV_F: FB_BIN(f_bits(as_f(y) / as_f(x))) break;
        case FB_X_LT_I: FB_BIN((uint64_t)((int64_t)x < (int64_t)y)) break;
        case FB_X_LE_I: FB_BIN((uint64_t)((int64_t)x <= (int64_t)y)) break;
        case FB_X_GT_I: FB_BIN((uint64_t)((int64_t)x > (int64_t)y)) break;
        case FB_X_GE_I: FB_BIN((uint64_t)((int64_t)x >= (int64_t)y)) break;
        case FB_X_EQ_I: FB_BIN((uint64_t)(x == y)) break;
        case FB_X_NE_I: FB_BIN((uint64_t)(x != y)) break;
        case FB_X_LT_F: FB_BIN((uint64_t)(as_f(x) < as_f(y))) break;
        case FB_X_LE_F: FB_BIN((uint64_t)(as_f(x) <= as_f(y))) break;
        case FB_X_GT_F: FB_BIN((uint64_t)(as_f(x) > as_f(y))) break;
        case FB_X_GE_F: FB_BIN((uint64_t)(as_f(x) >= as_f(y))) break;
        case FB_X_EQ_F: FB_BIN((uint64_t)(as_f(x) == as_f(y))) break;
        case FB_X_NE_F: FB_BIN((uint64_t)(as_f(x) != as_f(y))) break;
        case FB_X_AND: {  // Kleene: FALSE wins over NULL
          unsigned nv = 0;
          FB_ROWS(const bool va = (accv >> k) & 1u, vb = (bv >> k) & 1u;
                  const bool fa = va && acc[k] == 0, fb = vb && b[k] == 0; const bool isf = fa || fb;
                  if (isf || (va && vb)) nv |= 1u << k; acc[k] = (uint64_t)(!isf && va && vb);)
          accv = nv;
          break;
        }
        case FB_X_OR: {  // Kleene: TRUE wins over NULL
          unsigned nv = 0;
          FB_ROWS(const bool va = (accv >> k) & 1u, vb = (bv >> k) & 1u;
                  const bool ta = va && acc[k] != 0, tb = vb && b[k] != 0; const bool ist = ta || tb;
                  if (ist || (va && vb)) nv |= 1u << k; acc[k] = (uint64_t)ist;)
          accv = nv;
          break;
        }
        case FB_X_COALESCE: FB_ROWS(if (!((accv >> k) & 1u)) acc[k] = b[k];) accv |= bv; break;
        case FB_X_RCOALESCE: FB_ROWS(if ((bv >> k) & 1u) acc[k] = b[k];) accv |= bv; break;
        default: break;
      }
#undef FB_BIN
#undef FB_UN
#undef FB_ROWS
    }
  }
}

__global__ void __launch_bounds__(kExprThreads, 3)
fb_eval_expr_kernel(const __grid_constant__ ExprProgram P, int64_t nrows) {
  extern __shared__ __align__(16) uint64_t s_expr[];
  uint64_t* tmp_v = s_expr;                                              // [FB_EXPR_NREGS][kExprTile]
  uint8_t* tmp_m = (uint8_t*)(tmp_v + (size_t)FB_EXPR_NREGS * kExprTile);  // [FB_EXPR_NREGS][kExprThreads] bits
  const int tid = threadIdx.x;
  const int64_t ntiles = (nrows + kExprTile - 1) / kExprTile;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t row0 = tile * kExprTile;
    if (nrows - row0 >= kExprTile) run_tile<true>(P, row0, kExprTile, tid, tmp_v, tmp_m);
    else run_tile<false>(P, row0, (int)(nrows - row0), tid, tmp_v, tmp_m);
  }
}

constexpr size_t kExprSmem = (size_t)FB_EXPR_NREGS * kExprTile * 8 + (size_t)FB_EXPR_NREGS * kExprThreads;

}  // namespace

extern "C" int fb_eval_expr(int dev, void* stream, int64_t nrows, int ncols, const void* const* col_ptrs,
                            const int32_t* col_types, const uint8_t* const* col_valid, int nins,
                            const fb_expr_ins* program, int nouts, const int32_t* out_types,
                            void* const* out_ptrs, uint8_t* const* out_valid) {
  FB_CHECK(nrows >= 0, "nrows < 0");
  FB_CHECK(ncols >= 0 && ncols <= FB_EXPR_MAX_COLS, "ncols=%d out of range [0,%d]", ncols, FB_EXPR_MAX_COLS);
  FB_CHECK(nins >= 1 && nins <= FB_EXPR_MAX_INS, "nins=%d out of range [1,%d]", nins, FB_EXPR_MAX_INS);
  FB_CHECK(nouts >= 1 && nouts <= FB_EXPR_MAX_OUTS, "nouts=%d out of range [1,%d]", nouts, FB_EXPR_MAX_OUTS);
  FB_CHECK(program != nullptr && out_types != nullptr && out_ptrs != nullptr, "NULL argument");
  ExprProgram P;
  memset(&P, 0, sizeof(P));
  for (int c = 0; c < ncols; ++c) {
    FB_CHECK(col_types[c] >= FB_T_I8 && col_types[c] <= FB_T_F64, "column %d has unknown type %d", c, col_types[c]);
    FB_CHECK(nrows == 0 || col_ptrs[c] != nullptr, "column %d pointer is NULL", c);
    P.col_ptr[c] = col_ptrs[c];
    P.col_type[c] = col_types[c];
    P.col_valid[c] = col_valid ? col_valid[c] : nullptr;
  }
  for (int o = 0; o < nouts; ++o) {
    FB_CHECK(out_types[o] >= FB_T_I8 && out_types[o] <= FB_T_F64, "output %d: unknown type", o);
    FB_CHECK(nrows == 0 || out_ptrs[o] != nullptr, "output %d pointer is NULL", o);
    P.out_type[o] = out_types[o];
    P.out_ptr[o] = out_ptrs[o];
    P.out_valid[o] = out_valid ? out_valid[o] : nullptr;
  }
  for (int i = 0; i < nins; ++i) {
    const fb_expr_ins& in = program[i];
    FB_CHECK(in.op >= FB_X_MOV && in.op <= FB_X_RCOALESCE, "instruction %d: unknown op %d", i, in.op);
    FB_CHECK(in.kind >= FB_XK_NONE && in.kind <= FB_XK_NULL, "instruction %d: unknown operand kind %d", i, in.kind);
    if (in.op == FB_X_ST) {
      FB_CHECK(in.b >= 0 && in.b < FB_EXPR_NREGS, "instruction %d: temporary %d out of range", i, in.b);
      FB_CHECK(in.kind == FB_XK_NONE, "instruction %d: FB_X_ST takes no operand", i);
    } else if (in.op == FB_X_OUT) {
      FB_CHECK(in.b >= 0 && in.b < nouts, "instruction %d: output %d out of range", i, in.b);
      FB_CHECK(in.kind == FB_XK_NONE, "instruction %d: FB_X_OUT takes no operand", i);
    } else if (in.kind == FB_XK_REG) {
      FB_CHECK(in.b >= 0 && in.b < FB_EXPR_NREGS, "instruction %d: temporary %d out of range", i, in.b);
    } else if (in.kind == FB_XK_COL) {
      FB_CHECK(in.b >= 0 && in.b < ncols, "instruction %d: column %d out of range", i, in.b);
    }
    const bool binary = in.op == FB_X_MOV || in.op >= FB_X_ADD_I;
    FB_CHECK(!binary || in.kind != FB_XK_NONE, "instruction %d: op %d needs an operand", i, in.op);
    P.ins[i] = in;
  }
  P.nins = nins;
  if (nrows == 0) return 0;
  FbDeviceGuard guard(dev);
  FB_CHECK(guard.ok, "cannot select device %d", dev);
  static std::mutex mu;
  static uint64_t optin_done = 0;
  {
    std::lock_guard<std::mutex> lock(mu);
    if (!(dev >= 0 && dev < 64 && ((optin_done >> dev) & 1))) {
      FB_CUDA(cudaFuncSetAttribute(fb_eval_expr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)kExprSmem));
      if (dev >= 0 && dev < 64) optin_done |= 1ull << dev;
    }
  }
  const int64_t ntiles = (nrows + kExprTile - 1) / kExprTile;
  int64_t grid = (int64_t)fb_sm_count(dev) * 3;  // 65 KB of temporaries per CTA: 3 CTAs per SM
  if (grid > ntiles) grid = ntiles;
  fb_eval_expr_kernel<<<(unsigned)grid, kExprThreads, kExprSmem, (cudaStream_t)stream>>>(P, nrows);
  FB_CUDA(cudaGetLastError());
  return 0;
}
