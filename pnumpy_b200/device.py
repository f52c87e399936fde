"""Device-resident buffers and the atop ops on them, through the C ABI only.

``DeviceArray`` is a 1-D typed slab of HBM on one GPU.  The functions below call the same
launchers the NumPy hook layer calls (``pncu_Get*`` in ``include/pnumpy_cuda.h``); they exist so
that tests and benchmarks can exercise device-resident data (no PCIe in the timed region) the way
SURVEY.md 8d asks, and so that a user can keep intermediates on the GPU across a ufunc chain.
"""
import ctypes

import numpy as np

from . import cabi


class DeviceArray:
    """1-D contiguous array in device memory (current device at construction)."""

    def __init__(self, n, dtype, device=None):
        lib = cabi.load()
        cabi.init()
        if device is not None:
            cabi.check(lib.pncu_set_device(int(device)), "pncu_set_device")
        dev = ctypes.c_int(0)
        cabi.check(lib.pncu_get_device(ctypes.byref(dev)), "pncu_get_device")
        self.device = dev.value
        self.dtype = np.dtype(dtype)
        self.size = int(n)
        self.nbytes = self.size * self.dtype.itemsize
        p = ctypes.c_void_p()
        cabi.check(lib.pncu_malloc(ctypes.byref(p), max(self.nbytes, 16)), "pncu_malloc")
        self.ptr = p.value
        self._owner = True

    @property
    def atype(self):
        return cabi.atop_of(self.dtype)

    def __len__(self):
        return self.size

    def free(self):
        if getattr(self, "_owner", False) and self.ptr:
            lib = cabi.load()
            lib.pncu_set_device(self.device)
            lib.pncu_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def view(self, start, n):
        """A non-owning window [start, start+n) of this array."""
        v = object.__new__(DeviceArray)
        v.device, v.dtype, v.size = self.device, self.dtype, int(n)
        v.nbytes = v.size * self.dtype.itemsize
        v.ptr = self.ptr + int(start) * self.dtype.itemsize
        v._owner = False
        v._base = self
        return v

    def copy_from_host(self, a, stream=None):
        a = np.ascontiguousarray(a, dtype=self.dtype).reshape(-1)
        assert a.size == self.size
        lib = cabi.load()
        lib.pncu_set_device(self.device)
        cabi.check(lib.pncu_memcpy_h2d(self.ptr, a.ctypes.data, self.nbytes, stream), "h2d")
        cabi.check(lib.pncu_stream_sync(stream), "sync")
        return self

    def to_host(self, stream=None):
        out = np.empty(self.size, dtype=self.dtype)
        lib = cabi.load()
        lib.pncu_set_device(self.device)
        cabi.check(lib.pncu_memcpy_d2h(out.ctypes.data, self.ptr, self.nbytes, stream), "d2h")
        cabi.check(lib.pncu_stream_sync(stream), "sync")
        return out


def to_device(a, device=None):
    a = np.ascontiguousarray(a).reshape(-1)
    return DeviceArray(a.size, a.dtype, device).copy_from_host(a)


def empty(n, dtype, device=None):
    return DeviceArray(n, dtype, device)


def sync(stream=None):
    cabi.check(cabi.load().pncu_stream_sync(stream), "pncu_stream_sync")


def _ptr(x):
    return x.ptr if isinstance(x, DeviceArray) else x


def _stride(x, scalar):
    return 0 if scalar else x.dtype.itemsize


def binary(op, a, b, out=None, stream=None, a_scalar=False, b_scalar=False):
    """out = a <op> b with op a key of cabi.BINARY_OPS; scalars are 1-element DeviceArrays."""
    fn, out_t = cabi.get_binary(cabi.BINARY_OPS[op], a.atype)
    if fn is None:
        raise cabi.PnumpyCudaError(f"no kernel for {op} on {a.dtype}")
    n = b.size if a_scalar else a.size
    if out is None:
        out = DeviceArray(n, cabi.dtype_of(out_t), a.device)
    cabi.check(fn(a.ptr, b.ptr, out.ptr, n, _stride(a, a_scalar), _stride(b, b_scalar),
                  out.dtype.itemsize, stream), op)
    return out


def int_true_divide(a, b, out=None, stream=None):
    fn = cabi.get_int_true_divide(a.atype)
    if fn is None:
        raise cabi.PnumpyCudaError(f"no int true_divide kernel for {a.dtype}")
    if out is None:
        out = DeviceArray(a.size, np.float64, a.device)
    cabi.check(fn(a.ptr, b.ptr, out.ptr, a.size, a.dtype.itemsize, b.dtype.itemsize, 8, stream), "int_true_divide")
    return out


def muladd(a, b, c, out=None, stream=None):
    """out = a*b + c in one pass (two roundings; bit-identical to multiply then add)."""
    fn = cabi.get_muladd(a.atype)
    if fn is None:
        raise cabi.PnumpyCudaError(f"no fused multiply-add kernel for {a.dtype}")
    if out is None:
        out = DeviceArray(a.size, a.dtype, a.device)
    cabi.check(fn(a.ptr, b.ptr, c.ptr, out.ptr, a.size, stream), "muladd")
    return out


def muladd_sum(a, b, c, out=None, stream=None):
    """one-element DeviceArray holding sum(a*b + c), bit-identical to the three-call chain."""
    fn = cabi.get_muladd_sum(a.atype)
    if fn is None:
        raise cabi.PnumpyCudaError(f"no fused multiply-add-sum kernel for {a.dtype}")
    if out is None:
        out = DeviceArray(1, a.dtype, a.device)
    cabi.check(fn(a.ptr, b.ptr, c.ptr, out.ptr, a.size, stream), "muladd_sum")
    return out


def compare(op, a, b, out=None, stream=None, a_scalar=False, b_scalar=False):
    fn, _ = cabi.get_compare(cabi.COMP_OPS[op], a.atype)
    if fn is None:
        raise cabi.PnumpyCudaError(f"no kernel for compare {op} on {a.dtype}")
    n = b.size if a_scalar else a.size
    if out is None:
        out = DeviceArray(n, np.bool_, a.device)
    cabi.check(fn(a.ptr, b.ptr, out.ptr, n, _stride(a, a_scalar), _stride(b, b_scalar), 1, stream), op)
    return out


def unary(op, a, out=None, stream=None):
    if op in cabi.UNARY_OPS:
        fn, out_t = cabi.get_unary(cabi.UNARY_OPS[op], a.atype)
    else:
        fn, out_t = cabi.get_trig(cabi.TRIG_OPS[op], a.atype)
    if fn is None:
        raise cabi.PnumpyCudaError(f"no kernel for {op} on {a.dtype}")
    if out is None:
        out = DeviceArray(a.size, cabi.dtype_of(out_t), a.device)
    cabi.check(fn(a.ptr, out.ptr, a.size, a.dtype.itemsize, out.dtype.itemsize, stream), op)
    return out


def astype(a, dtype, out=None, stream=None):
    """out = a converted to `dtype` with NumPy's cast semantics (pncu_GetConvertFast)."""
    dtype = np.dtype(dtype)
    fn = cabi.get_convert(a.atype, cabi.atop_of(dtype))
    if fn is None:
        raise cabi.PnumpyCudaError(f"no conversion kernel {a.dtype} -> {dtype}")
    if out is None:
        out = DeviceArray(a.size, dtype, a.device)
    cabi.check(fn(a.ptr, out.ptr, a.size, a.dtype.itemsize, out.dtype.itemsize, stream), "astype")
    return out


def reduce(op, a, out=None, start=None, stream=None):
    """*out = fold(op, a) [op *start]; returns the 1-element DeviceArray `out`."""
    fn = cabi.get_reduce(cabi.BINARY_OPS[op], a.atype)
    if fn is None:
        raise cabi.PnumpyCudaError(f"no reduction kernel for {op} on {a.dtype}")
    if out is None:
        out = DeviceArray(1, a.dtype, a.device)
    cabi.check(fn(a.ptr, out.ptr, _ptr(start) if start is not None else None, a.size,
                  a.dtype.itemsize, stream), "reduce " + op)
    return out


def argminmax(kind, a, out=None, stream=None):
    fn = cabi.get_argminmax(kind, a.atype)
    if fn is None:
        raise cabi.PnumpyCudaError(f"no arg-reduction kernel on {a.dtype}")
    if out is None:
        out = DeviceArray(1, np.int64, a.device)
    cabi.check(fn(a.ptr, a.size, out.ptr, stream), "argminmax")
    return out


class Timer:
    """CUDA-event timer on a stream of the C ABI (torch.cuda.Event would not see this stream)."""

    def __init__(self):
        lib = cabi.load()
        self.a, self.b = ctypes.c_void_p(), ctypes.c_void_p()
        cabi.check(lib.pncu_event_create(ctypes.byref(self.a)), "event")
        cabi.check(lib.pncu_event_create(ctypes.byref(self.b)), "event")

    def start(self, stream=None):
        cabi.check(cabi.load().pncu_event_record(self.a, stream), "record")

    def stop(self, stream=None):
        cabi.check(cabi.load().pncu_event_record(self.b, stream), "record")

    def ms(self):
        lib = cabi.load()
        cabi.check(lib.pncu_event_sync(self.b), "event sync")
        ms = ctypes.c_float(0)
        cabi.check(lib.pncu_event_elapsed_ms(self.a, self.b, ctypes.byref(ms)), "elapsed")
        return ms.value
