"""Randomised end-to-end stress of the hooked path against NumPy's own loops: random ops, dtypes, operand kinds
(pageable / pinned / strided views / in place), sizes up to well past one staging ring per lane (so slots are reused
many times), random lane counts.  Every result is compared bit for bit.   python tools/stress.py [seconds]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("PNUMPY_B200_AUTOINIT", "0")
import pnumpy_b200 as pn  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.default_rng(2024)
pn.init()
pn.enable()
pn.cuda_setthreshold(0)
BIG = 1 << 27   # elements of the largest case: 512 MiB of float32, 1 GiB of float64
pool = {}
for dt in ("int16", "int32", "float32", "float64"):
    base = rng.integers(-1000, 1000, BIG + 1024).astype(dt) if dt.startswith("int") else rng.normal(0, 100, BIG + 1024).astype(dt)
    pool[dt] = (base, np.ascontiguousarray(base[::-1]), pn.pinned_array(base), pn.pinned_array(base[::-1]))
binary = [np.add, np.subtract, np.multiply, np.maximum, np.greater, np.less_equal]
unary = [np.abs, np.negative]
t_end = time.time() + budget
done = 0
while time.time() < t_end:
    dt = rng.choice(list(pool))
    pa, pb, qa, qb = pool[dt]
    lg = rng.choice([10, 14, 16, 17, 20, 22, 24, 25, 26, 27], p=[.1, .1, .1, .1, .15, .15, .1, .08, .07, .05])
    n = int((1 << lg) + rng.integers(-5, 300))
    n = max(1, min(n, BIG))
    off = int(rng.integers(0, 64))
    kind = rng.choice(["pageable", "pinned", "mixed", "strided", "inplace"])
    workers = int(rng.choice([1, 2, 3, 7]))
    pn.thread_setworkers(workers)
    if kind == "pageable":
        a, b = pa[off:off + n], pb[off:off + n]
    elif kind == "pinned":
        a, b = qa[off:off + n], qb[off:off + n]
    elif kind == "mixed":
        a, b = qa[off:off + n], pb[off:off + n]
    elif kind == "strided":
        m = min(n, BIG // 3)
        a, b = pa[off:off + 3 * m:3], pb[off + 2 * m:off:-2][:m]
        n = min(a.size, b.size)
        a, b = a[:n], b[:n]
    else:
        a, b = pa[off:off + n].copy(), pb[off:off + n]
    f = binary[int(rng.integers(len(binary)))] if rng.random() < 0.8 else unary[int(rng.integers(len(unary)))]
    with np.errstate(all="ignore"):
        if f in unary:
            pn.enable(); got = f(a)
            pn.disable(); want = f(a)
        elif kind == "inplace" and f not in (np.greater, np.less_equal):
            a2 = a.copy()
            pn.enable(); f(a, b, out=a); got = a
            pn.disable(); f(a2, b, out=a2); want = a2
        else:
            pn.enable(); got = f(a, b)
            pn.disable(); want = f(a, b)
        pn.enable(); r_got = (a.max(), a.min(), int(a.argmax())) if n else None
        pn.disable(); r_want = (a.max(), a.min(), int(a.argmax())) if n else None
    pn.enable()
    if got.tobytes() != want.tobytes() or r_got != r_want:
        print("MISMATCH", dt, n, off, kind, workers, f.__name__)
        sys.exit(1)
    done += 1
print(f"stress ok: {done} random cases in {budget:.0f} s, GPU calls {pn.cuda_info()['calls']}")
