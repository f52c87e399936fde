"""Pinned host ndarrays: operands the nte dispatcher can DMA in place (no staging copy).

``recycler_enable()`` makes every large ndarray pinned automatically; these helpers do it for one
array at a time."""
import ctypes

import numpy as np

from . import _pnumpy

__all__ = ['pinned_empty', 'pinned_array']


class _PinnedBlock:
    """Owner of one cudaHostAlloc block; exposes the buffer protocol through ctypes."""

    def __init__(self, nbytes):
        self.capsule, self.address = _pnumpy.pinned_alloc(int(nbytes))
        self.nbytes = int(nbytes)

    def buffer(self):
        return (ctypes.c_char * max(self.nbytes, 1)).from_address(self.address)


def pinned_empty(shape, dtype=np.float64):
    """``np.empty`` in page-locked host memory (freed when the array and its views die)."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) if not np.isscalar(shape) else int(shape)
    block = _PinnedBlock(n * dtype.itemsize)
    buf = block.buffer()
    buf._pnumpy_owner = block  # keeps the capsule (and the allocation) alive with the buffer
    a = np.frombuffer(buf, dtype=dtype, count=n)
    return a.reshape(shape)


def pinned_array(a):
    """A pinned copy of ``a``."""
    a = np.asarray(a)
    out = pinned_empty(a.shape, a.dtype)
    out[...] = a
    return out
