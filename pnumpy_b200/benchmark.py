"""``pn.benchmark()`` plumbing: ratio of plain NumPy time to pnumpy time per dtype.

reference: ``src/pnumpy/benchmark.py:13-174`` -- same data (``arange(n) % 253 + 1``), same protocol
(one dry run, then the median of 100 ``timer_gettsc`` deltas with ``out=`` recycled), same
off-then-on switching, same CSV rows.  The only addition is ``pinned=``: allocate the operands in
pinned host memory so the GPU path DMAs them in place.
"""
import numpy as np

from ._pnumpy import atop_enable, atop_disable
from ._pnumpy import thread_enable, thread_disable
from ._pnumpy import timer_gettsc
from .pinned import pinned_array, pinned_empty

__all__ = ['benchmark', 'benchmark_func', 'benchmark_timeit']

_DEFAULT_CTYPES = [np.bool_, np.int8, np.int16, np.int32, np.int64, np.float32, np.float64]


def benchmark_timeit(func=np.equal, ctypes=_DEFAULT_CTYPES, scalar=False, unary=False, reduct=False,
                     outdtype=None, recycle=True, sizes=[1_000_000], pinned=False, loop_size=100):
    """Internal routine to benchmark a function: median rdtsc ticks per dtype."""
    timedelta = np.zeros(len(ctypes), np.int64)
    for s in sizes:
        for slot, ctype in enumerate(ctypes):
            if ctype is np.bool_:
                a = np.arange(s, dtype=np.int8).astype(ctype) + 1
            else:
                # the reference's data, `arange(s) % 253 + 1` (benchmark.py:58-61), built in int64 and cast: NumPy 2
                # refuses `int8_array % 253` (NEP 50), where NumPy < 1.20 silently widened the array
                a = (np.arange(s, dtype=np.int64) % 253 + 1).astype(ctype)
            if pinned:
                a = pinned_array(a)
            b = a[5] if scalar else (pinned_array(a) if pinned else a.copy())

            def time_func(recycle, c):
                start = timer_gettsc()
                if not unary:
                    result = func(a, b, out=c) if recycle else func(a, b)
                elif reduct:
                    result = func(a)
                else:
                    result = func(a, out=c) if recycle else func(a)
                return timer_gettsc() - start, result

            # dry run (also produces the recycled output array)
            _, c = time_func(False, None)
            if pinned and isinstance(c, np.ndarray) and c.ndim:
                c = pinned_empty(c.shape, c.dtype)
            ticks = np.zeros(loop_size, np.int64)
            for loop in range(loop_size):
                delta, result = time_func(recycle, c)
                del result
                ticks[loop] = delta
            timedelta[slot] = np.median(ticks)
    return timedelta


def benchmark_func(func=np.equal, ctypes=_DEFAULT_CTYPES, scalar=False, unary=False, reduct=False,
                   outdtype=None, recycle=True, atop=True, thread=True, sizes=[1_000_000], pinned=False,
                   loop_size=100):
    """
    Benchmark one function: t(numpy loops) / t(pnumpy loops) per dtype; > 1 is an improvement.

    Examples
    --------
    benchmark_func(np.add)
    benchmark_func(np.add, sizes=[2**16])
    benchmark_func(np.sqrt, unary=True)
    """
    kw = dict(func=func, ctypes=ctypes, scalar=scalar, unary=unary, reduct=reduct, outdtype=outdtype,
              recycle=recycle, sizes=sizes, pinned=pinned, loop_size=loop_size)
    # disable atop and threading, get the original time
    atop_disable()
    thread_disable()
    t0 = benchmark_timeit(**kw)
    # now possibly enable atop and threading
    if atop:
        atop_enable()
    if thread:
        thread_enable()
    t1 = benchmark_timeit(**kw)
    return t0 / t1


def benchmark(ctypes=_DEFAULT_CTYPES, recycle=True, atop=True, thread=True, sizes=[1_000_000], pinned=False,
              loop_size=100):
    """
    Performs a simple benchmark of the ratio of normal numpy vs pnumpy.  The output is formatted to
    be copied and pasted in a csv file.  A result above 1.0 indicates an improvement.

    Examples
    --------
    pn.benchmark()
    pn.benchmark(thread=False)
    pn.benchmark(sizes=[2**16])
    pn.benchmark(ctypes=[np.float32, np.float64])
    """
    def header(ct):
        return f'{sizes[0]} rows,' + ''.join(f'{i.__name__},' for i in ct)

    def row(name, data):
        print(f'{name},' + ''.join(f'{i:5.2f},' for i in data))

    kw = dict(ctypes=ctypes, recycle=recycle, atop=atop, thread=thread, sizes=sizes, pinned=pinned, loop_size=loop_size)
    print(header(ctypes))
    row("a==b", benchmark_func(np.equal, scalar=False, unary=False, outdtype='?', **kw))
    row("a==5", benchmark_func(np.equal, scalar=True, unary=False, outdtype='?', **kw))
    row("a+b", benchmark_func(np.add, scalar=False, unary=False, **kw))
    row("a+5", benchmark_func(np.add, scalar=True, unary=False, **kw))
    row("a/5", benchmark_func(np.true_divide, scalar=True, unary=False, **kw))
    row("abs", benchmark_func(np.abs, scalar=False, unary=True, **kw))
    row("isnan", benchmark_func(np.isnan, scalar=False, unary=True, **kw))
    row("sin", benchmark_func(np.sin, scalar=False, unary=True, **kw))
    row("log", benchmark_func(np.log, scalar=False, unary=True, **kw))
    row("sum", benchmark_func(np.sum, scalar=False, reduct=True, unary=True, **kw))
    row("min", benchmark_func(np.min, scalar=False, reduct=True, unary=True, **kw))
