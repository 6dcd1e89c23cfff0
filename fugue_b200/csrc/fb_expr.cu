// Vectorised column-expression evaluator for sm_100a.
//
// ExecutionEngine.select / filter / assign (fugue/execution/execution_engine.py:736-887) evaluate
// expression trees (fugue/column/expressions.py) row by row; the reference turns them into SQL text and
// hands them to qpd/pandas, which materialises one temporary column per operator.  Here a whole
// SELECT list (+ WHERE predicate) is compiled on the host into one short register-machine program and
// evaluated in ONE pass over the input columns:
//
//   * a CTA owns a tile of 1024 rows (256 threads x 4 rows); the machine's vector registers live in
//     shared memory ([reg][1024] 8-byte values + [reg][1024] validity bytes), each thread only ever
//     touches its own 4 lanes of every register, so instructions need no barriers;
//   * the interpreter dispatches once per instruction per thread and then runs the 4 rows, so the
//     switch is amortised and the kernel stays HBM-bound: every referenced column is read once,
//     every output written once, no temporaries in global memory;
//   * values are canonical 64-bit (int64 or float64, bool as int64 0/1); the host compiler types
//     every instruction, inserts the int->float conversions and allocates registers;
//   * SQL NULL semantics: arithmetic and comparisons propagate NULL, AND/OR are Kleene three-valued,
//     IS NULL / IS NOT NULL / COALESCE read the validity lane.
#include <mutex>

#include "fb_common.cuh"

namespace {

constexpr int kExprThreads = 256;
constexpr int kExprItems = 4;
constexpr int kExprTile = kExprThreads * kExprItems;

struct ExprProgram {
  const void* col_ptr[FB_EXPR_MAX_COLS];
  const uint8_t* col_valid[FB_EXPR_MAX_COLS];
  int32_t col_type[FB_EXPR_MAX_COLS];
  void* out_ptr[FB_EXPR_MAX_OUTS];
  uint8_t* out_valid[FB_EXPR_MAX_OUTS];
  int32_t out_reg[FB_EXPR_MAX_OUTS];
  int32_t out_type[FB_EXPR_MAX_OUTS];
  int32_t nins, nouts;
  fb_expr_ins ins[FB_EXPR_MAX_INS];
};

__device__ __forceinline__ uint64_t load_as_bits(const void* p, int32_t type, int64_t row) {
  switch (type) {
    case FB_T_I8: return (uint64_t)(int64_t)((const int8_t*)p)[row];
    case FB_T_I16: return (uint64_t)(int64_t)((const int16_t*)p)[row];
    case FB_T_I32: return (uint64_t)(int64_t)((const int32_t*)p)[row];
    case FB_T_I64: return (uint64_t)((const int64_t*)p)[row];
    case FB_T_U8: return (uint64_t)((const uint8_t*)p)[row];
    case FB_T_F32: return (uint64_t)__double_as_longlong((double)((const float*)p)[row]);
    default: return (uint64_t)((const int64_t*)p)[row];  // FB_T_F64: raw bits
  }
}

__device__ __forceinline__ void store_from_bits(void* p, int32_t type, int64_t row, uint64_t bits) {
  switch (type) {
    case FB_T_I8: ((int8_t*)p)[row] = (int8_t)(int64_t)bits; break;
    case FB_T_I16: ((int16_t*)p)[row] = (int16_t)(int64_t)bits; break;
    case FB_T_I32: ((int32_t*)p)[row] = (int32_t)(int64_t)bits; break;
    case FB_T_I64: ((int64_t*)p)[row] = (int64_t)bits; break;
    case FB_T_U8: ((uint8_t*)p)[row] = (uint8_t)bits; break;
    case FB_T_F32: ((float*)p)[row] = (float)__longlong_as_double((long long)bits); break;
    default: ((int64_t*)p)[row] = (int64_t)bits; break;  // FB_T_F64
  }
}

__device__ __forceinline__ double as_f(uint64_t b) { return __longlong_as_double((long long)b); }
__device__ __forceinline__ uint64_t f_bits(double d) { return (uint64_t)__double_as_longlong(d); }

__global__ void __launch_bounds__(kExprThreads, 3)
fb_eval_expr_kernel(const __grid_constant__ ExprProgram P, int64_t nrows) {
  extern __shared__ __align__(16) uint64_t s_expr[];
  uint64_t* vals = s_expr;                                           // [FB_EXPR_NREGS][kExprTile]
  uint8_t* valid = (uint8_t*)(vals + (size_t)FB_EXPR_NREGS * kExprTile);  // [FB_EXPR_NREGS][kExprTile]
  const int tid = threadIdx.x;
  const int64_t ntiles = (nrows + kExprTile - 1) / kExprTile;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t row0 = tile * kExprTile;
    const int nk = (int)(nrows - row0 < kExprTile ? nrows - row0 : kExprTile);  // rows in this tile
    for (int pc = 0; pc < P.nins; ++pc) {
      const fb_expr_ins in = P.ins[pc];
      uint64_t* dv = vals + (size_t)in.dst * kExprTile;
      uint8_t* dm = valid + (size_t)in.dst * kExprTile;
      const uint64_t* av = vals + (size_t)in.a * kExprTile;
      const uint8_t* am = valid + (size_t)in.a * kExprTile;
      const uint64_t* bv = vals + (size_t)in.b * kExprTile;
      const uint8_t* bm = valid + (size_t)in.b * kExprTile;
#define FB_EACH(...)                                  \
  _Pragma("unroll") for (int k = 0; k < kExprItems; ++k) { \
    const int i = k * kExprThreads + tid;             \
    if (i < nk) { __VA_ARGS__ }                       \
  }
#define FB_BIN(EXPR) FB_EACH(const uint64_t x = av[i], y = bv[i]; dv[i] = (EXPR); dm[i] = am[i] & bm[i];)
#define FB_UN(EXPR) FB_EACH(const uint64_t x = av[i]; dv[i] = (EXPR); dm[i] = am[i];)
      switch (in.op) {
        case FB_X_LOAD: {
          const void* p = P.col_ptr[in.a];
          const uint8_t* m = P.col_valid[in.a];
          const int32_t t = P.col_type[in.a];
          FB_EACH(dv[i] = load_as_bits(p, t, row0 + i); dm[i] = m ? (uint8_t)(m[row0 + i] != 0) : (uint8_t)1;)
          break;
        }
        case FB_X_LIT: FB_EACH(dv[i] = (uint64_t)in.imm; dm[i] = 1;) break;
        case FB_X_NULL: FB_EACH(dv[i] = 0; dm[i] = 0;) break;
        case FB_X_MOV: FB_UN(x) break;
        case FB_X_I2F: FB_UN(f_bits((double)(int64_t)x)) break;
        case FB_X_F2I: FB_UN((uint64_t)(int64_t)as_f(x)) break;
        case FB_X_ADD_I: FB_BIN(x + y) break;
        case FB_X_SUB_I: FB_BIN(x - y) break;
        case FB_X_MUL_I: FB_BIN(x * y) break;
        case FB_X_NEG_I: FB_UN(0 - x) break;
        case FB_X_ADD_F: FB_BIN(f_bits(as_f(x) + as_f(y))) break;
        case FB_X_SUB_F: FB_BIN(f_bits(as_f(x) - as_f(y))) break;
        case FB_X_MUL_F: FB_BIN(f_bits(as_f(x) * as_f(y))) break;
        case FB_X_DIV_F: FB_BIN(f_bits(as_f(x) / as_f(y))) break;
        case FB_X_NEG_F: FB_UN(x ^ 0x8000000000000000ull) break;
        case FB_X_LT_I: FB_BIN((uint64_t)((int64_t)x < (int64_t)y)) break;
        case FB_X_LE_I: FB_BIN((uint64_t)((int64_t)x <= (int64_t)y)) break;
        case FB_X_EQ_I: FB_BIN((uint64_t)(x == y)) break;
        case FB_X_NE_I: FB_BIN((uint64_t)(x != y)) break;
        case FB_X_LT_F: FB_BIN((uint64_t)(as_f(x) < as_f(y))) break;
        case FB_X_LE_F: FB_BIN((uint64_t)(as_f(x) <= as_f(y))) break;
        case FB_X_EQ_F: FB_BIN((uint64_t)(as_f(x) == as_f(y))) break;
        case FB_X_NE_F: FB_BIN((uint64_t)(as_f(x) != as_f(y))) break;
        case FB_X_AND:  // Kleene: FALSE wins over NULL
          FB_EACH(const bool va = am[i], vb = bm[i]; const bool fa = va && av[i] == 0, fb = vb && bv[i] == 0;
                  const bool isf = fa || fb; dm[i] = (uint8_t)(isf || (va && vb));
                  dv[i] = (uint64_t)(!isf && va && vb);)
          break;
        case FB_X_OR:  // Kleene: TRUE wins over NULL
          FB_EACH(const bool va = am[i], vb = bm[i]; const bool ta = va && av[i] != 0, tb = vb && bv[i] != 0;
                  const bool ist = ta || tb; dm[i] = (uint8_t)(ist || (va && vb)); dv[i] = (uint64_t)ist;)
          break;
        case FB_X_NOT: FB_UN((uint64_t)(x == 0)) break;
        case FB_X_IS_NULL: FB_EACH(dv[i] = (uint64_t)(am[i] == 0); dm[i] = 1;) break;
        case FB_X_NOT_NULL: FB_EACH(dv[i] = (uint64_t)(am[i] != 0); dm[i] = 1;) break;
        case FB_X_COALESCE:
          FB_EACH(const bool va = am[i]; dv[i] = va ? av[i] : bv[i]; dm[i] = (uint8_t)(va | bm[i]);)
          break;
        case FB_X_TOBOOL_I: FB_UN((uint64_t)(x != 0)) break;
        case FB_X_TOBOOL_F: FB_UN((uint64_t)(as_f(x) != 0.0)) break;
        default: break;
      }
#undef FB_BIN
#undef FB_UN
    }
    for (int o = 0; o < P.nouts; ++o) {
      const uint64_t* __restrict__ rv = vals + (size_t)P.out_reg[o] * kExprTile;
      const uint8_t* __restrict__ rm = valid + (size_t)P.out_reg[o] * kExprTile;
      void* op = P.out_ptr[o];
      uint8_t* om = P.out_valid[o];
      const int32_t t = P.out_type[o];
      FB_EACH(store_from_bits(op, t, row0 + i, rm[i] ? rv[i] : 0ull); if (om) om[row0 + i] = rm[i];)
    }
#undef FB_EACH
  }
}

constexpr size_t kExprSmem = (size_t)FB_EXPR_NREGS * kExprTile * 9;

}  // namespace

extern "C" int fb_eval_expr(int dev, void* stream, int64_t nrows, int ncols, const void* const* col_ptrs,
                            const int32_t* col_types, const uint8_t* const* col_valid, int nins,
                            const fb_expr_ins* program, int nouts, const int32_t* out_regs,
                            const int32_t* out_types, void* const* out_ptrs, uint8_t* const* out_valid) {
  FB_CHECK(nrows >= 0, "nrows < 0");
  FB_CHECK(ncols >= 0 && ncols <= FB_EXPR_MAX_COLS, "ncols=%d out of range [0,%d]", ncols, FB_EXPR_MAX_COLS);
  FB_CHECK(nins >= 1 && nins <= FB_EXPR_MAX_INS, "nins=%d out of range [1,%d]", nins, FB_EXPR_MAX_INS);
  FB_CHECK(nouts >= 1 && nouts <= FB_EXPR_MAX_OUTS, "nouts=%d out of range [1,%d]", nouts, FB_EXPR_MAX_OUTS);
  FB_CHECK(program != nullptr && out_regs != nullptr && out_types != nullptr && out_ptrs != nullptr,
           "NULL argument");
  ExprProgram P;
  memset(&P, 0, sizeof(P));
  for (int c = 0; c < ncols; ++c) {
    FB_CHECK(col_types[c] >= FB_T_I8 && col_types[c] <= FB_T_F64, "column %d has unknown type %d", c, col_types[c]);
    FB_CHECK(nrows == 0 || col_ptrs[c] != nullptr, "column %d pointer is NULL", c);
    P.col_ptr[c] = col_ptrs[c];
    P.col_type[c] = col_types[c];
    P.col_valid[c] = col_valid ? col_valid[c] : nullptr;
  }
  for (int i = 0; i < nins; ++i) {
    const fb_expr_ins& in = program[i];
    FB_CHECK(in.op >= FB_X_LOAD && in.op <= FB_X_TOBOOL_F, "instruction %d: unknown op %d", i, in.op);
    FB_CHECK(in.dst >= 0 && in.dst < FB_EXPR_NREGS, "instruction %d: dst register %d out of range", i, in.dst);
    if (in.op == FB_X_LOAD) {
      FB_CHECK(in.a >= 0 && in.a < ncols, "instruction %d: column %d out of range", i, in.a);
    } else {
      FB_CHECK(in.a >= 0 && in.a < FB_EXPR_NREGS && in.b >= 0 && in.b < FB_EXPR_NREGS,
               "instruction %d: source register out of range", i);
    }
    P.ins[i] = in;
    if (in.op == FB_X_LOAD) P.ins[i].b = 0;  // the kernel forms (unused) register pointers from a and b
  }
  for (int o = 0; o < nouts; ++o) {
    FB_CHECK(out_regs[o] >= 0 && out_regs[o] < FB_EXPR_NREGS, "output %d: register out of range", o);
    FB_CHECK(out_types[o] >= FB_T_I8 && out_types[o] <= FB_T_F64, "output %d: unknown type", o);
    FB_CHECK(nrows == 0 || out_ptrs[o] != nullptr, "output %d pointer is NULL", o);
    P.out_reg[o] = out_regs[o];
    P.out_type[o] = out_types[o];
    P.out_ptr[o] = out_ptrs[o];
    P.out_valid[o] = out_valid ? out_valid[o] : nullptr;
  }
  P.nins = nins;
  P.nouts = nouts;
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
  int64_t grid = (int64_t)fb_sm_count(dev) * 3;  // 72 KB of registers per CTA: 3 CTAs per SM
  if (grid > ntiles) grid = ntiles;
  fb_eval_expr_kernel<<<(unsigned)grid, kExprThreads, kExprSmem, (cudaStream_t)stream>>>(P, nrows);
  FB_CUDA(cudaGetLastError());
  return 0;
}
