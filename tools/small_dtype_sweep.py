"""Device-resident GB/s of the small dtypes (bool, int8, uint8, int16, uint16) on 2^28 elements."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("PNUMPY_B200_AUTOINIT", "0")
from pnumpy_b200 import cabi, device as dev
from tools.perf_sweep import time_call
cabi.init(); lib = cabi.load()
n = 1 << 28
rng = np.random.default_rng(1)
for dt in ("bool", "int8", "uint8", "int16", "uint16"):
    s = np.dtype(dt).itemsize
    piece = (rng.integers(0, 2, 1 << 22) if dt == "bool" else rng.integers(1, 120, 1 << 22)).astype(dt)
    a, b, o, ob = dev.empty(n, dt), dev.empty(n, dt), dev.empty(n, dt), dev.empty(n, np.bool_)
    r1, ri = dev.empty(1, dt), dev.empty(1, np.int64)   # (results preallocated: a cudaMalloc per call would be 10 us of the timing)
    dp = dev.to_device(piece)
    for x in (a, b):
        for off in range(0, n, piece.size):
            lib.pncu_memcpy_d2d(x.ptr + off * s, dp.ptr, piece.nbytes, None)
    dev.sync()
    cases = [("add", 3 * s, lambda: dev.binary("ADD", a, b, out=o)), ("greater", 2 * s + 1, lambda: dev.compare("GT", a, b, out=ob)),
             ("abs", 2 * s, lambda: dev.unary("ABS", a, out=o)), ("isnan", s + 1, lambda: dev.unary("ISNAN", a, out=ob)),
             ("sum", s, lambda: dev.reduce("ADD", a, out=r1)), ("min", s, lambda: dev.reduce("MIN", a, out=r1)), ("max", s, lambda: dev.reduce("MAX", a, out=r1)),
             ("argmax", s, lambda: dev.argminmax(0, a, out=ri)), ("argmin", s, lambda: dev.argminmax(1, a, out=ri))]
    for name, bpe, fn in cases:
        avg, best = time_call(fn)
        print(f"{name:8s} {dt:7s} {avg:8.4f} ms {bpe * n / avg / 1e6:8.1f} GB/s {100 * bpe * n / avg / 1e6 / 8000:5.1f}% of 8 TB/s", flush=True)
    for x in (a, b, o, ob): x.free()
