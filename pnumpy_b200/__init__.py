"""
pnumpy_b200 -- pnumpy's atop ufunc hot path on NVIDIA B200 GPUs, behind pnumpy's own hooks.

Like pnumpy (reference: ``src/pnumpy/__init__.py``), importing the package calls ``init`` to set
it up: NumPy's ufunc inner loops (add/subtract/multiply/divide, the comparisons, abs/isnan/sqrt/...,
sin/cos/log/exp and the add/min/max reductions, argmin/argmax) are replaced with wrapped versions
through ``PyUFunc_ReplaceLoopBySignature``.  You can then enable/disable the subsystems:

  - atop
        The fast inner loops.  Here they are hand-written sm_100a CUDA kernels in
        ``libpnumpy_cuda`` (``include/pnumpy_cuda.h``); ``atop_disable()`` routes every call to
        NumPy's original loop, as in the reference.

  - threading
        The reference splits a loop into 16K-element chunks over worker threads.  Here the ``nte``
        dispatcher (``include/pnumpy_nte.h``) splits a call into contiguous shards over *lanes*:
        lane ``l`` drives GPU ``l % G`` through its own H2D / compute / D2H streams and pinned
        staging ring.  ``thread_setworkers(n)`` sets the number of extra lanes (the reference's
        extra threads); ``thread_disable()`` pins every call to one lane on GPU 0.

  - ledger
        While enabled: per-category call counters (GPU loop vs NumPy loop) and a ring of per-call
        records (op, dtype, elements, bytes, start/end ticks, where it ran): ``ledger_records()``.

  - recycler
        A NEP-49 allocator.  ``recycler_enable()``: large ndarrays are born in recycled *pinned*
        host memory, so the dispatcher can DMA them in place instead of staging them.
        ``recycler_enable(kind="managed")``: they are born in CUDA *managed* memory instead -- the
        arrays NumPy itself creates (``np.empty``, ufunc outputs, the temporaries of ``a*b+c``) are
        device-capable, hooked calls on them run in place on the GPU at every size, and nothing
        crosses PCIe until host code touches a result.

There is no CPU fallback inside the GPU library: importing this package on a machine without a
B200 raises ImportError (the reference raises on a CPU without AVX2).  Set
``PNUMPY_B200_AUTOINIT=0`` to import without initialising (tools, CPU-side tests).
"""
import os as _os

__version__ = "0.1.0"

__all__ = [
    'initialize', 'atop_enable', 'atop_disable', 'atop_isenabled', 'atop_info', 'atop_setworkers', 'cpustring',
    'thread_enable', 'thread_disable', 'thread_isenabled', 'thread_getworkers', 'thread_setworkers', 'thread_zigzag',
    'ledger_enable', 'ledger_disable', 'ledger_isenabled', 'ledger_info', 'ledger_records', 'ledger_clear',
    'recycler_enable', 'recycler_disable', 'recycler_isenabled', 'recycler_info', 'recycler_trim',
    'timer_gettsc', 'timer_getutc', 'benchmark', 'benchmark_func', 'init', 'enable', 'disable', 'cpu_count_linux',
    'argmin', 'argmax',
    # additions of this build (nothing above changes meaning)
    'cuda_info', 'cuda_reset_stats', 'cuda_setthreshold', 'cuda_getthreshold', 'cuda_setslab', 'cuda_inject_fault',
    'pinned_empty', 'pinned_array', 'managed_empty', 'managed_array', 'muladd', 'muladd_sum', 'astype',
]

import numpy as np  # noqa: E402

try:
    from . import _pnumpy
except ImportError as _e:  # the extension is the product: no Python fallback
    raise ImportError(
        "pnumpy_b200._pnumpy (the CUDA hook extension) is not built: run "
        "`python -c 'import __graft_entry__ as g; g.build()'` or `make -C pnumpy_b200/csrc`") from _e

from ._pnumpy import atop_enable, atop_disable, atop_isenabled, atop_info, atop_setworkers, cpustring  # noqa: E402
from ._pnumpy import thread_enable, thread_disable, thread_isenabled, thread_getworkers, thread_setworkers, thread_zigzag  # noqa: E402
from ._pnumpy import timer_gettsc, timer_getutc  # noqa: E402
from ._pnumpy import ledger_enable, ledger_disable, ledger_isenabled, ledger_info, ledger_records, ledger_clear  # noqa: E402
from ._pnumpy import recycler_enable, recycler_disable, recycler_isenabled, recycler_info, recycler_trim  # noqa: E402
from ._pnumpy import cuda_info, cuda_reset_stats, cuda_setthreshold, cuda_getthreshold, cuda_setslab, cuda_inject_fault  # noqa: E402

from .cpu import cpu_count_linux, init, enable, disable  # noqa: E402
from .benchmark import benchmark, benchmark_func  # noqa: E402
from .pinned import pinned_empty, pinned_array, managed_empty, managed_array  # noqa: E402


def initialize():
    """Kept for compatibility with pnumpy (``src/pnumpy/__init__.py:70-72``): same as ``init()``."""
    init()


def argmax(a, axis=None, out=None):
    """Indices of the maximum values along an axis (``ndarray.argmax``; first occurrence, first NaN
    wins).  reference: ``src/pnumpy/sort.py:410-483`` -- a thin forward to NumPy, whose per-dtype
    ``argmax`` slot this package has replaced with the GPU arg-reduction."""
    return np.asanyarray(a).argmax(axis=axis, out=out)


def argmin(a, axis=None, out=None):
    """Indices of the minimum values along an axis.  reference: ``src/pnumpy/sort.py:486-563``."""
    return np.asanyarray(a).argmin(axis=axis, out=out)


def muladd(a, b, c, out=None):
    """``a*b + c`` in one pass (the reference's "for longer equations: a=b*c+d" case).

    Same values, bit for bit, as ``np.add(np.multiply(a, b), c, out=out)``: the product and the sum
    are rounded separately, there is no FMA.  For equal-shape C-contiguous arrays of one dtype
    (int32/uint32/int64/uint64/float32/float64) above the dispatch threshold the operands cross PCIe
    / HBM once and no temporary is materialised; anything else runs the two hooked ufunc calls."""
    a = np.asanyarray(a)
    if out is None and isinstance(a, np.ndarray):
        from .pinned import pinned_empty as _pe, managed_empty as _me
        from ._pnumpy import pointer_kind as _kind
        k = _kind(a.ctypes.data) if a.size else 0
        res = _me(a.shape, a.dtype) if k >= 2 else (_pe(a.shape, a.dtype) if k == 1 else np.empty_like(a))
    else:
        res = out
    if isinstance(res, np.ndarray) and _pnumpy.fused_muladd(a, b, c, res):
        return res
    return np.add(np.multiply(a, b), c, out=out)


def astype(a, dtype, out=None):
    """``a.astype(dtype)`` on the GPU -- dtype conversion is the first item of pnumpy's roadmap
    (``doc_src/source/roadmap.rst:11-15``); NumPy has no public hook for its cast loops, hence a function.

    Same values as NumPy's cast for bool, the integer types, float32 and float64: integers wrap, float -> int
    truncates toward zero (and yields what x86's conversion instructions yield when the value does not fit,
    which is what NumPy returns, with a warning), int -> float and float64 -> float32 round to nearest even.
    The result lives in the same kind of memory as ``a`` (managed stays managed, pinned stays pinned);
    anything the GPU path does not take (other dtypes, non-contiguous arrays, small host arrays) is NumPy's
    own ``astype``."""
    a = np.asanyarray(a)
    dtype = np.dtype(dtype)
    if out is None:
        k = _pnumpy.pointer_kind(a.ctypes.data) if a.size else 0
        res = managed_empty(a.shape, dtype) if k >= 2 else (pinned_empty(a.shape, dtype) if k == 1 else np.empty(a.shape, dtype))
    else:
        res = out
        if res.dtype != dtype or res.shape != a.shape:
            raise ValueError("out must have the requested dtype and the shape of a")
    if _pnumpy.convert(a, res):
        return res
    if out is None:
        return a.astype(dtype)
    np.copyto(out, a, casting="unsafe")
    return out


def muladd_sum(a, b, c):
    """``(a*b + c).sum()`` in one pass: 12 bytes per float32 element instead of 28, no temporaries.

    Bit-identical to ``np.add(np.multiply(a, b), c).sum()`` as this package computes it (same 64 KiB
    blocks, same fold order, float32 accumulated in float64); falls back to exactly that chain when
    the operands do not qualify (see ``muladd``)."""
    r = _pnumpy.fused_muladd_sum(a, b, c)
    if r is not None:
        return r
    return np.add(np.multiply(a, b), c).sum()


# start the engine by default (reference: src/pnumpy/__init__.py:74-76)
if _os.environ.get("PNUMPY_B200_AUTOINIT", "1") != "0":
    init()
