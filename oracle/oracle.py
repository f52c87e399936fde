"""TEST INFRASTRUCTURE -- ctypes drivers for the CPU oracle (``libatop_oracle.so``, built from
``atop_oracle.c``) and, when present, for the reference's own loop bodies
(``_ref/libatop_ref.so``, compiled unmodified from /root/reference/src/atop).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  Nothing under pnumpy_b200/ does.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libatop_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libatop_ref.so")
REF_EXT_DIR = os.path.join(HERE, "_ref")

_ATOP = {"bool": 0, "int8": 1, "uint8": 2, "int16": 3, "uint16": 4, "int32": 5, "uint32": 6,
         "int64": 7, "uint64": 8, "float32": 12, "float64": 13}
BINARY = dict(ADD=1, SUB=2, MUL=3, MIN=5, MAX=6, FLOORDIV=9, DIV=13, LOGICAL_AND=16, LOGICAL_OR=18,
              BITWISE_LSHIFT=19, BITWISE_RSHIFT=20, BITWISE_AND=21, BITWISE_XOR=22, BITWISE_OR=23)
COMPARE = dict(EQ=0, NE=1, LT=2, GT=3, LTE=4, GTE=5)
UNARY = dict(ABS=1, SIGNBIT=2, INVERT=4, FLOOR=5, CEIL=6, TRUNC=7, ROUND=8, NEGATIVE=9, SIGN=11,
             SQRT=15, LOGICAL_NOT=18, ISINF=19, ISNAN=20, ISFINITE=21, BITWISE_NOT=28)
TRIG = dict(SIN=1, COS=2, TAN=3, ASIN=4, ACOS=5, ATAN=6, SINH=7, COSH=8, TANH=9, ASINH=10, ACOSH=11, ATANH=12,
            LOG=13, LOG2=14, LOG10=15, EXP=16, EXP2=17, EXPM1=18, LOG1P=19, CBRT=20)

_i64 = ctypes.c_int64
_vp = ctypes.c_void_p


def atop_of(dtype):
    return _ATOP[np.dtype(dtype).name]


def build():
    """(Re)build libatop_oracle.so -- and the _ref artefacts when /root/reference is present."""
    subprocess.check_call(["make", "-s", "-C", HERE])


_olib = None


def lib():
    global _olib
    if _olib is None:
        if not os.path.exists(ORACLE_SO):
            build()
        L = ctypes.CDLL(ORACLE_SO)
        L.oracle_binary.argtypes = [ctypes.c_int, ctypes.c_int, _vp, _vp, _vp, _i64, _i64, _i64, _i64]
        L.oracle_compare.argtypes = L.oracle_binary.argtypes
        L.oracle_int_true_divide.argtypes = [ctypes.c_int, _vp, _vp, _vp, _i64, _i64, _i64, _i64]
        L.oracle_unary.argtypes = [ctypes.c_int, ctypes.c_int, _vp, _vp, _i64, _i64, _i64]
        L.oracle_trig.argtypes = L.oracle_unary.argtypes
        L.oracle_reduce.argtypes = [ctypes.c_int, ctypes.c_int, _vp, _vp, _vp, _i64, _i64]
        L.oracle_argminmax.argtypes = [ctypes.c_int, ctypes.c_int, _vp, _i64, ctypes.POINTER(_i64)]
        L.oracle_convert.argtypes = [ctypes.c_int, ctypes.c_int, _vp, _vp, _i64]
        _olib = L
    return _olib


def _st(a):
    """(pointer, byte stride) of a 1-D array or a 0-d/1-element scalar (stride 0)."""
    a = np.asarray(a)
    if a.ndim == 0 or (a.ndim == 1 and a.size == 1 and a.strides[0] == 0):
        return a.ctypes.data, 0
    assert a.ndim == 1
    return a.ctypes.data, a.strides[0]


def _n(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return max(a.size if a.ndim else 1, b.size if b.ndim else 1)


def binary(op, a, b, out=None):
    a, b = np.asarray(a), np.asarray(b)
    dt = a.dtype
    n = _n(a, b)
    if out is None:
        out = np.empty(n, dtype=dt)
    (pa, sa), (pb, sb) = _st(a), _st(b)
    rc = lib().oracle_binary(BINARY[op], atop_of(dt), pa, pb, out.ctypes.data, n, sa, sb, out.strides[0])
    if rc:
        raise NotImplementedError(f"oracle: {op} {dt}")
    return out


def int_true_divide(a, b):
    a, b = np.asarray(a), np.asarray(b)
    n = _n(a, b)
    out = np.empty(n, dtype=np.float64)
    (pa, sa), (pb, sb) = _st(a), _st(b)
    rc = lib().oracle_int_true_divide(atop_of(a.dtype), pa, pb, out.ctypes.data, n, sa, sb, 8)
    if rc:
        raise NotImplementedError(f"oracle: int_true_divide {a.dtype}")
    return out


def compare(op, a, b):
    a, b = np.asarray(a), np.asarray(b)
    n = _n(a, b)
    out = np.empty(n, dtype=np.bool_)
    (pa, sa), (pb, sb) = _st(a), _st(b)
    rc = lib().oracle_compare(COMPARE[op], atop_of(a.dtype), pa, pb, out.ctypes.data, n, sa, sb, 1)
    if rc:
        raise NotImplementedError(f"oracle: compare {op} {a.dtype}")
    return out


_BOOL_OUT = {"SIGNBIT", "LOGICAL_NOT", "ISINF", "ISNAN", "ISFINITE"}


def unary(op, a):
    a = np.asarray(a)
    if op in TRIG:
        out = np.empty(a.size, dtype=a.dtype)
        rc = lib().oracle_trig(TRIG[op], atop_of(a.dtype), a.ctypes.data, out.ctypes.data, a.size,
                               a.strides[0], out.strides[0])
    else:
        out = np.empty(a.size, dtype=np.bool_ if op in _BOOL_OUT else a.dtype)
        rc = lib().oracle_unary(UNARY[op], atop_of(a.dtype), a.ctypes.data, out.ctypes.data, a.size,
                                a.strides[0], out.strides[0])
    if rc:
        raise NotImplementedError(f"oracle: {op} {a.dtype}")
    return out


def astype(a, dtype):
    """a.astype(dtype) as plain C casts (oracle_convert)"""
    a = np.ascontiguousarray(a).reshape(-1)
    out = np.empty(a.size, dtype=np.dtype(dtype))
    rc = lib().oracle_convert(atop_of(a.dtype), atop_of(out.dtype), a.ctypes.data, out.ctypes.data, a.size)
    if rc:
        raise NotImplementedError(f"oracle: convert {a.dtype} -> {out.dtype}")
    return out


def reduce(op, a, start=None):
    a = np.asarray(a)
    out = np.zeros(1, dtype=a.dtype)
    st = None
    if start is not None:
        st = np.asarray([start], dtype=a.dtype)
    rc = lib().oracle_reduce(BINARY[op], atop_of(a.dtype), a.ctypes.data, out.ctypes.data,
                             st.ctypes.data if st is not None else None, a.size,
                             a.strides[0] if a.size else a.dtype.itemsize)
    if rc:
        raise NotImplementedError(f"oracle: reduce {op} {a.dtype}")
    return out[0]


def argminmax(kind, a):
    a = np.ascontiguousarray(a)
    idx = _i64(0)
    rc = lib().oracle_argminmax(kind, atop_of(a.dtype), a.ctypes.data, a.size, ctypes.byref(idx))
    if rc:
        raise NotImplementedError(f"oracle: argminmax {a.dtype}")
    return idx.value


# ---- the reference's own loop bodies (only where oracle/_ref was built) -------------------------
_UNARY_FN = ctypes.CFUNCTYPE(None, _vp, _vp, _i64, _i64, _i64)
_TWO_FN = ctypes.CFUNCTYPE(None, _vp, _vp, _vp, _i64, _i64, _i64, _i64)
_RED_FN = ctypes.CFUNCTYPE(None, _vp, _vp, _vp, _i64, _i64)
_rlib = None


def have_ref():
    return os.path.exists(REF_SO)


def ref():
    """libatop_ref.so: extern "C" getters of src/atop/atop.h:19-29."""
    global _rlib
    if _rlib is None:
        L = ctypes.CDLL(REF_SO)
        L.atop_init()
        for name in ("GetSimpleMathOpFast", "GetComparisonOpFast"):
            f = getattr(L, name)
            f.restype = _vp
            f.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
        for name in ("GetUnaryOpFast", "GetUnaryOpSlow", "GetTrigOpFast"):
            f = getattr(L, name)
            f.restype = _vp
            f.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
        L.GetReduceMathOpFast.restype = _vp
        L.GetReduceMathOpFast.argtypes = [ctypes.c_int, ctypes.c_int]
        _rlib = L
    return _rlib


def ref_binary(op, a, b, compare_op=False):
    """Run the reference's fast loop body; returns None when it has none for (op, dtype)."""
    a, b = np.asarray(a), np.asarray(b)
    o = ctypes.c_int(-1)
    t = atop_of(a.dtype)
    if compare_op:
        p = ref().GetComparisonOpFast(COMPARE[op], t, t, ctypes.byref(o))
    else:
        p = ref().GetSimpleMathOpFast(BINARY[op], t, t, ctypes.byref(o))
    if not p:
        return None
    n = _n(a, b)
    out = np.empty(n, dtype=np.bool_ if compare_op else a.dtype)
    (pa, sa), (pb, sb) = _st(a), _st(b)
    _TWO_FN(p)(pa, pb, out.ctypes.data, n, sa, sb, out.strides[0])
    return out


def ref_unary(op, a):
    a = np.asarray(a)
    o = ctypes.c_int(-1)
    t = atop_of(a.dtype)
    if op in TRIG:
        p = ref().GetTrigOpFast(TRIG[op], t, ctypes.byref(o))
    else:
        p = ref().GetUnaryOpFast(UNARY[op], t, ctypes.byref(o)) or ref().GetUnaryOpSlow(UNARY[op], t, ctypes.byref(o))
    if not p:
        return None
    out = np.empty(a.size, dtype=np.bool_ if o.value == 0 else a.dtype)
    _UNARY_FN(p)(a.ctypes.data, out.ctypes.data, a.size, a.strides[0], out.strides[0])
    return out


def ref_reduce(op, a, start):
    """The reference's reduce body over the whole array with an explicit start value."""
    a = np.asarray(a)
    p = ref().GetReduceMathOpFast(BINARY[op], atop_of(a.dtype))
    if not p:
        return None
    out = np.zeros(1, dtype=a.dtype)
    st = np.asarray([start], dtype=a.dtype)
    _RED_FN(p)(a.ctypes.data, out.ctypes.data, st.ctypes.data, a.size, a.strides[0])
    return out[0]


def ref_getter_table():
    """Out-types and fast-pointer presence of every reference getter (see tests/golden)."""
    L = ref()
    types = list(range(0, 9)) + [12, 13, 14]

    def two(f, op):
        row = {}
        for t in types:
            o = ctypes.c_int(-1)
            p = f(op, t, t, ctypes.byref(o))
            row[str(t)] = [o.value, bool(p)]
        return row

    def unary_row(op):
        row = {}
        for t in types:
            o = ctypes.c_int(-1)
            p = L.GetUnaryOpFast(op, t, ctypes.byref(o))
            if not p:
                p = L.GetUnaryOpSlow(op, t, ctypes.byref(o))
            row[str(t)] = [o.value, bool(p)]
        return row

    def trig_row(op):
        row = {}
        for t in types:
            o = ctypes.c_int(-1)
            p = L.GetTrigOpFast(op, t, ctypes.byref(o))
            row[str(t)] = [o.value, bool(p)]
        return row

    return {
        "binary": {str(op): two(L.GetSimpleMathOpFast, op) for op in range(1, 29)},
        "compare": {str(op): two(L.GetComparisonOpFast, op) for op in range(0, 6)},
        "unary": {str(op): unary_row(op) for op in range(1, 29)},
        "trig": {str(op): trig_row(op) for op in range(1, 21)},
        "reduce": {str(op): {str(t): bool(L.GetReduceMathOpFast(op, t)) for t in types} for op in (1, 3, 5, 6)},
    }
