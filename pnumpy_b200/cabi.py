"""ctypes binding of ``libpnumpy_cuda.so`` (declared in ``include/pnumpy_cuda.h``).

This is the reference-side stub a maintainer would write to reach the CUDA loop bodies from
Python without the CPython extension; the extension (``pnumpy_b200._pnumpy``) links the same
library directly.  There is no fallback: if the shared library is missing or no sm_100 device
is present the calls raise.

The enum values are the reference's (``src/atop/common_inc.h:236-372``).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PNUMPY_CUDA_LIB") or os.path.join(_HERE, "lib", "libpnumpy_cuda.so")  # env: tuning builds only

# ---- vocabulary (reference: src/atop/common_inc.h:236-372) -------------------------------------
ATOP_BOOL, ATOP_INT8, ATOP_UINT8, ATOP_INT16, ATOP_UINT16 = 0, 1, 2, 3, 4
ATOP_INT32, ATOP_UINT32, ATOP_INT64, ATOP_UINT64 = 5, 6, 7, 8
ATOP_FLOAT, ATOP_DOUBLE, ATOP_LONGDOUBLE = 12, 13, 14

BINARY_OPS = dict(ADD=1, SUB=2, MUL=3, MOD=4, MIN=5, MAX=6, NANMIN=7, NANMAX=8, FLOORDIV=9,
                  POWER=10, REMAINDER=11, FMOD=12, DIV=13, LOGICAL_AND=16, LOGICAL_XOR=17,
                  LOGICAL_OR=18, BITWISE_LSHIFT=19, BITWISE_RSHIFT=20, BITWISE_AND=21,
                  BITWISE_XOR=22, BITWISE_OR=23)
COMP_OPS = dict(EQ=0, NE=1, LT=2, GT=3, LTE=4, GTE=5)
UNARY_OPS = dict(ABS=1, SIGNBIT=2, FABS=3, INVERT=4, FLOOR=5, CEIL=6, TRUNC=7, ROUND=8,
                 NEGATIVE=9, POSITIVE=10, SIGN=11, RINT=12, SQRT=15, SQUARE=16, RECIPROCAL=17,
                 LOGICAL_NOT=18, ISINF=19, ISNAN=20, ISFINITE=21, ISNORMAL=22, ISNOTINF=23,
                 ISNOTNAN=24, ISNOTFINITE=25, ISNOTNORMAL=26, ISNANORZERO=27, BITWISE_NOT=28)
TRIG_OPS = dict(SIN=1, COS=2, TAN=3, ASIN=4, ACOS=5, ATAN=6, SINH=7, COSH=8, TANH=9, ASINH=10,
                ACOSH=11, ATANH=12, LOG=13, LOG2=14, LOG10=15, EXP=16, EXP2=17, EXPM1=18,
                LOG1P=19, CBRT=20)
ARGMAX, ARGMIN = 0, 1

# numpy dtype char -> atop type (reference: convert_dtype_to_atop, src/pnumpy/_pnumpy.cpp:7-27)
_DTYPE_TO_ATOP = {"bool": 0, "int8": 1, "uint8": 2, "int16": 3, "uint16": 4, "int32": 5,
                  "uint32": 6, "int64": 7, "uint64": 8, "float32": 12, "float64": 13}
_ATOP_TO_DTYPE = {v: k for k, v in _DTYPE_TO_ATOP.items()}


def atop_of(dtype):
    import numpy as np
    return _DTYPE_TO_ATOP[np.dtype(dtype).name]


def dtype_of(atype):
    import numpy as np
    return np.dtype(_ATOP_TO_DTYPE[atype])


class PnumpyCudaError(RuntimeError):
    pass


_c_i64 = ctypes.c_int64
_vp = ctypes.c_void_p

UNARY_FN = ctypes.CFUNCTYPE(ctypes.c_int, _vp, _vp, _c_i64, _c_i64, _c_i64, _vp)
ANY_TWO_FN = ctypes.CFUNCTYPE(ctypes.c_int, _vp, _vp, _vp, _c_i64, _c_i64, _c_i64, _c_i64, _vp)
REDUCE_FN = ctypes.CFUNCTYPE(ctypes.c_int, _vp, _vp, _vp, _c_i64, _c_i64, _vp)
ARGMINMAX_FN = ctypes.CFUNCTYPE(ctypes.c_int, _vp, _c_i64, _vp, _vp)
ANY_THREE_FN = ctypes.CFUNCTYPE(ctypes.c_int, _vp, _vp, _vp, _vp, _c_i64, _vp)   # also pncu_three_reduce_fn

# every symbol include/pnumpy_cuda.h declares: name -> (restype, argtypes)
_P_INT = ctypes.POINTER(ctypes.c_int)
SIGNATURES = {
    "pncu_init": (ctypes.c_int, [_P_INT]),
    "pncu_shutdown": (ctypes.c_int, []),
    "pncu_device_count": (ctypes.c_int, []),
    "pncu_device_info": (ctypes.c_int, [ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t]),
    "pncu_last_error": (ctypes.c_char_p, []),
    "pncu_launch_count": (_c_i64, []),
    "pncu_set_tunable": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_int]),
    "pncu_GetSimpleMathOpFast": (_vp, [ctypes.c_int, ctypes.c_int, ctypes.c_int, _P_INT]),
    "pncu_GetReduceMathOpFast": (_vp, [ctypes.c_int, ctypes.c_int]),
    "pncu_GetComparisonOpFast": (_vp, [ctypes.c_int, ctypes.c_int, ctypes.c_int, _P_INT]),
    "pncu_GetUnaryOpFast": (_vp, [ctypes.c_int, ctypes.c_int, _P_INT]),
    "pncu_GetTrigOpFast": (_vp, [ctypes.c_int, ctypes.c_int, _P_INT]),
    "pncu_GetArgMinMaxOpFast": (_vp, [ctypes.c_int, ctypes.c_int]),
    "pncu_GetIntTrueDivide": (_vp, [ctypes.c_int]),
    "pncu_GetMulAddFast": (_vp, [ctypes.c_int]),
    "pncu_GetMulAddSumFast": (_vp, [ctypes.c_int]),
    "pncu_GetConvertFast": (_vp, [ctypes.c_int, ctypes.c_int]),
    "pncu_muladd_sum_partials": (ctypes.c_int, [ctypes.c_int, _vp, _vp, _vp, _c_i64, _vp, _vp]),
    "pncu_itemsize": (ctypes.c_int, [ctypes.c_int]),
    "pncu_reduce_block_elems": (_c_i64, []),
    "pncu_reduce_group_slots": (ctypes.c_int, []),
    "pncu_last_tail_ns": (_c_i64, []),
    "pncu_reduce_partial_count": (_c_i64, [ctypes.c_int, _c_i64]),
    "pncu_reduce_partial_bytes": (ctypes.c_int, [ctypes.c_int, ctypes.c_int]),
    "pncu_reduce_partials": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, _vp, _c_i64, _vp, _vp]),
    "pncu_reduce_finish": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, _vp, _c_i64, _vp, _vp, _vp]),
    "pncu_argminmax_partials": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, _vp, _c_i64, _c_i64, _vp, _vp]),
    "pncu_argminmax_finish": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, _vp, _c_i64, _vp, _vp]),
    "pncu_set_device": (ctypes.c_int, [ctypes.c_int]),
    "pncu_get_device": (ctypes.c_int, [_P_INT]),
    "pncu_malloc": (ctypes.c_int, [ctypes.POINTER(_vp), ctypes.c_size_t]),
    "pncu_free": (ctypes.c_int, [_vp]),
    "pncu_malloc_host": (ctypes.c_int, [ctypes.POINTER(_vp), ctypes.c_size_t]),
    "pncu_free_host": (ctypes.c_int, [_vp]),
    "pncu_host_register": (ctypes.c_int, [_vp, ctypes.c_size_t]),
    "pncu_host_unregister": (ctypes.c_int, [_vp]),
    "pncu_memcpy_h2d": (ctypes.c_int, [_vp, _vp, ctypes.c_size_t, _vp]),
    "pncu_memcpy_d2h": (ctypes.c_int, [_vp, _vp, ctypes.c_size_t, _vp]),
    "pncu_memcpy_d2d": (ctypes.c_int, [_vp, _vp, ctypes.c_size_t, _vp]),
    "pncu_memset": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_size_t, _vp]),
    "pncu_stream_create": (ctypes.c_int, [ctypes.POINTER(_vp)]),
    "pncu_stream_destroy": (ctypes.c_int, [_vp]),
    "pncu_stream_sync": (ctypes.c_int, [_vp]),
    "pncu_device_sync": (ctypes.c_int, []),
    "pncu_event_create": (ctypes.c_int, [ctypes.POINTER(_vp)]),
    "pncu_event_destroy": (ctypes.c_int, [_vp]),
    "pncu_event_record": (ctypes.c_int, [_vp, _vp]),
    "pncu_event_sync": (ctypes.c_int, [_vp]),
    "pncu_event_elapsed_ms": (ctypes.c_int, [_vp, _vp, ctypes.POINTER(ctypes.c_float)]),
    "pncu_pointer_kind": (ctypes.c_int, [_vp, _P_INT]),
    "pncu_exchange_slices": (ctypes.c_int, []),
    "pncu_reduce_exchange_finish": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, _vp, _c_i64, _c_i64, _vp, ctypes.c_int, _c_i64, _vp, _vp,
                                                   ctypes.c_uint64, _vp, _vp]),
    "pncu_argminmax_exchange_finish": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, _vp, _c_i64, _c_i64, _vp, ctypes.c_int, _c_i64, _vp,
                                                      ctypes.c_uint64, _vp, _vp]),
    "pncu_exchange_timeouts": (_c_i64, []),
    "pncu_reduce_fused": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, _vp, _c_i64, _vp, _vp, _vp, _vp]),
    "pncu_argminmax_fused": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, _vp, _c_i64, _c_i64, _vp, _vp, _vp]),
    "pncu_muladd_sum_fused": (ctypes.c_int, [ctypes.c_int, _vp, _vp, _vp, _c_i64, _vp, _vp, _vp]),
    "pncu_fused_reset": (ctypes.c_int, [_vp]),
    "pncu_enable_peer_access": (ctypes.c_int, [ctypes.c_int]),
    "pncu_ipc_export": (ctypes.c_int, [_vp, _vp]),
    "pncu_ipc_import": (ctypes.c_int, [_vp, ctypes.POINTER(_vp)]),
    "pncu_ipc_close": (ctypes.c_int, [_vp]),
}

PNCU_MAX_PEERS = 8


class PeerTable(ctypes.Structure):
    """pncu_peer_table (include/pnumpy_cuda.h)"""
    _fields_ = [("n", ctypes.c_int), ("slots", _vp * PNCU_MAX_PEERS), ("arrived", _vp * PNCU_MAX_PEERS)]



class FusedExchange(ctypes.Structure):
    """pncu_fused_exchange (include/pnumpy_cuda.h)"""
    _fields_ = [("n", ctypes.c_int), ("self", ctypes.c_int), ("all", ctypes.c_int), ("root", ctypes.c_int),
                ("slot_base", _c_i64), ("total_count", _c_i64),
                ("slots", _vp * PNCU_MAX_PEERS), ("arrived", _vp * PNCU_MAX_PEERS),
                ("expected", ctypes.c_uint64), ("done_flag", _vp), ("done_value", ctypes.c_uint64),
                ("group_offset", _c_i64)]


FUSED_TIMEOUT_BIT = 1 << 63
ERR_TIMEOUT = -6

_lib = None


def load():
    """Load libpnumpy_cuda.so (no GPU needed for loading; first CUDA use needs a B200)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PnumpyCudaError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C pnumpy_b200/csrc`; pnumpy_b200 has no CPU fallback")
    lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().pncu_last_error().decode(errors="replace")
        raise PnumpyCudaError(f"{what} failed with status {rc}: {msg}")


def init():
    """pncu_init(); returns the device count or raises when no sm_100 device is usable."""
    n = ctypes.c_int(0)
    check(load().pncu_init(ctypes.byref(n)), "pncu_init")
    return n.value


def _getter(name, fntype, *args, with_out=True):
    lib = load()
    if with_out:
        out = ctypes.c_int(-1)
        p = getattr(lib, name)(*args, ctypes.byref(out))
        return (fntype(p) if p else None), out.value
    p = getattr(lib, name)(*args)
    return fntype(p) if p else None


def get_binary(op, atype):
    return _getter("pncu_GetSimpleMathOpFast", ANY_TWO_FN, op, atype, atype)


def get_compare(op, atype):
    return _getter("pncu_GetComparisonOpFast", ANY_TWO_FN, op, atype, atype)


def get_unary(op, atype):
    return _getter("pncu_GetUnaryOpFast", UNARY_FN, op, atype)


def get_trig(op, atype):
    return _getter("pncu_GetTrigOpFast", UNARY_FN, op, atype)


def get_reduce(op, atype):
    return _getter("pncu_GetReduceMathOpFast", REDUCE_FN, op, atype, with_out=False)


def get_argminmax(kind, atype):
    return _getter("pncu_GetArgMinMaxOpFast", ARGMINMAX_FN, kind, atype, with_out=False)


def get_int_true_divide(atype):
    return _getter("pncu_GetIntTrueDivide", ANY_TWO_FN, atype, with_out=False)


def get_muladd(atype):
    return _getter("pncu_GetMulAddFast", ANY_THREE_FN, atype, with_out=False)


def get_convert(in_atype, out_atype):
    return _getter("pncu_GetConvertFast", UNARY_FN, in_atype, out_atype, with_out=False)


def get_muladd_sum(atype):
    return _getter("pncu_GetMulAddSumFast", ANY_THREE_FN, atype, with_out=False)
