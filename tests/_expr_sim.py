"""numpy model of the fb_eval_expr accumulator machine (include/fugue_b200.h, K8): lets the host-side
expression compiler be checked against the oracle without a GPU.  Test infrastructure only."""
import numpy as np

from fugue_b200 import kernels as K

_NP_OF_T = {K.T_I8: np.int8, K.T_I16: np.int16, K.T_I32: np.int32, K.T_I64: np.int64, K.T_U8: np.uint8,
            K.T_F32: np.float32, K.T_F64: np.float64}


def _to_bits(a: np.ndarray) -> np.ndarray:
    if a.dtype.kind == "f":
        return a.astype(np.float64).view(np.uint64)
    return a.astype(np.int64).view(np.uint64)


def _f(b):
    return b.view(np.float64)


def _fb(x):
    return np.asarray(x, dtype=np.float64).view(np.uint64)


def _ib(x):
    return np.asarray(x).astype(np.int64).view(np.uint64)


def run(n, cols, valid, program, out_types):
    """cols: numpy arrays (storage dtype); valid: uint8 arrays or None.  Returns ([values], [valid])."""
    acc = np.zeros(n, dtype=np.uint64)
    accv = np.ones(n, dtype=bool)
    tmp = {}
    outs = [None] * len(out_types)
    outv = [None] * len(out_types)
    with np.errstate(all="ignore"):
        for op, kind, b, flags, imm in program:
            bv = np.ones(n, dtype=bool)
            bb = np.full(n, imm & ((1 << 64) - 1), dtype=np.uint64)
            if kind == K.XK_COL:
                bb = _to_bits(cols[b])
                if valid[b] is not None:
                    bv = valid[b] != 0
            elif kind == K.XK_REG:
                bb, bv = tmp[b]
            elif kind == K.XK_NULL:
                bv = np.zeros(n, dtype=bool)
            if flags & K.XF_B_I2F:
                bb = _fb(bb.view(np.int64).astype(np.float64))
            x, y = acc, bb
            xi, yi = x.view(np.int64), y.view(np.int64)
            if op == K.X_MOV:
                acc, accv = bb.copy(), bv.copy()
            elif op == K.X_ST:
                tmp[b] = (acc.copy(), accv.copy())
            elif op == K.X_OUT:
                t = out_types[b]
                vals = np.where(accv, acc, np.uint64(0))
                if t in (K.T_F32, K.T_F64):
                    outs[b] = _f(vals).astype(_NP_OF_T[t])
                else:
                    outs[b] = vals.view(np.int64).astype(_NP_OF_T[t])
                outv[b] = accv.astype(np.uint8)
            elif op == K.X_I2F:
                acc = _fb(xi.astype(np.float64))
            elif op == K.X_F2I:
                acc = _ib(np.trunc(np.nan_to_num(_f(x), nan=0.0, posinf=0.0, neginf=0.0)))
            elif op == K.X_NEG_I:
                acc = _ib(-xi)
            elif op == K.X_NEG_F:
                acc = _fb(-_f(x))
            elif op == K.X_NOT:
                acc = _ib(x == 0)
            elif op == K.X_IS_NULL:
                acc, accv = _ib(~accv), np.ones(n, dtype=bool)
            elif op == K.X_NOT_NULL:
                acc, accv = _ib(accv), np.ones(n, dtype=bool)
            elif op == K.X_TOBOOL_I:
                acc = _ib(x != 0)
            elif op == K.X_TOBOOL_F:
                acc = _ib(_f(x) != 0.0)
            elif op in (K.X_AND, K.X_OR):
                if op == K.X_AND:
                    fa, fb = accv & (x == 0), bv & (y == 0)
                    isf = fa | fb
                    nv = isf | (accv & bv)
                    acc = _ib(~isf & accv & bv)
                else:
                    ta, tb = accv & (x != 0), bv & (y != 0)
                    ist = ta | tb
                    nv = ist | (accv & bv)
                    acc = _ib(ist)
                accv = nv
            elif op == K.X_COALESCE:
                acc = np.where(accv, x, y)
                accv = accv | bv
            elif op == K.X_RCOALESCE:
                acc = np.where(bv, y, x)
                accv = accv | bv
            else:
                xf, yf = _f(x), _f(y)
                table = {
                    K.X_ADD_I: lambda: _ib(xi + yi), K.X_SUB_I: lambda: _ib(xi - yi), K.X_RSUB_I: lambda: _ib(yi - xi),
                    K.X_MUL_I: lambda: _ib(xi * yi),
                    K.X_ADD_F: lambda: _fb(xf + yf), K.X_SUB_F: lambda: _fb(xf - yf), K.X_RSUB_F: lambda: _fb(yf - xf),
                    K.X_MUL_F: lambda: _fb(xf * yf), K.X_DIV_F: lambda: _fb(xf / yf), K.X_RDIV_F: lambda: _fb(yf / xf),
                    K.X_LT_I: lambda: _ib(xi < yi), K.X_LE_I: lambda: _ib(xi <= yi), K.X_GT_I: lambda: _ib(xi > yi),
                    K.X_GE_I: lambda: _ib(xi >= yi), K.X_EQ_I: lambda: _ib(xi == yi), K.X_NE_I: lambda: _ib(xi != yi),
                    K.X_LT_F: lambda: _ib(xf < yf), K.X_LE_F: lambda: _ib(xf <= yf), K.X_GT_F: lambda: _ib(xf > yf),
                    K.X_GE_F: lambda: _ib(xf >= yf), K.X_EQ_F: lambda: _ib(xf == yf), K.X_NE_F: lambda: _ib(xf != yf),
                }
                acc = table[op]()
                accv = accv & bv
    return outs, outv
