"""ndarrays in memory the GPU can reach without a staging copy.

* ``pinned_empty`` / ``pinned_array``: page-locked HOST memory.  The nte dispatcher DMAs such operands
  in place (no memcpy through staging slabs).  ``recycler_enable()`` makes every large ndarray pinned
  automatically; these helpers do it for one array at a time.
* ``managed_empty`` / ``managed_array``: CUDA managed memory -- a DEVICE-RESIDENT ndarray.  NumPy and
  the kernels share one pointer.  A hooked ufunc whose array operands are all managed runs in place on
  the GPU at HBM speed (the dispatcher prefetches the pages to the GPU first); nothing crosses PCIe
  until host code reads the result, at which point the driver migrates the touched pages back.
"""
import ctypes

import numpy as np

from . import _pnumpy

__all__ = ['pinned_empty', 'pinned_array', 'managed_empty', 'managed_array']


class _Block:
    """Owner of one allocation; exposes the buffer protocol through ctypes."""

    def __init__(self, nbytes, alloc):
        self.capsule, self.address = alloc(int(nbytes))
        self.nbytes = int(nbytes)

    def buffer(self):
        return (ctypes.c_char * max(self.nbytes, 1)).from_address(self.address)


def _empty(shape, dtype, alloc):
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) if not np.isscalar(shape) else int(shape)
    block = _Block(n * dtype.itemsize, alloc)
    buf = block.buffer()
    buf._pnumpy_owner = block  # keeps the capsule (and the allocation) alive with the buffer
    return np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)


def pinned_empty(shape, dtype=np.float64):
    """``np.empty`` in page-locked host memory (freed when the array and its views die)."""
    return _empty(shape, dtype, _pnumpy.pinned_alloc)


def pinned_array(a):
    """A pinned copy of ``a``."""
    a = np.asarray(a)
    out = pinned_empty(a.shape, a.dtype)
    out[...] = a
    return out


def managed_empty(shape, dtype=np.float64):
    """``np.empty`` in CUDA managed memory: a device-resident ndarray (see the module docstring)."""
    return _empty(shape, dtype, _pnumpy.managed_alloc)


def managed_array(a):
    """A managed (device-resident) copy of ``a``."""
    a = np.asarray(a)
    out = managed_empty(a.shape, a.dtype)
    out[...] = a
    return out
