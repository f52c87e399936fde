"""Contiguous operands whose 32-byte alignments disagree (a[1:] - a[:-1], a[1:] + b[:-1], sqrt(a[3:])) on
device-resident 2^28-element arrays: CUDA-event GB/s of the realigning vector kernels (ew.cuh: ld_pack_realigned)
and of reductions over a misaligned start (ops_reduce.cu: load_block), with a bit-exact check against NumPy on a
prefix.   python tools/misaligned_probe.py [log2_elems]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("PNUMPY_B200_AUTOINIT", "0")
from pnumpy_b200 import cabi, device as dev  # noqa: E402

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 28
n = 1 << lg
cabi.init()
lib = cabi.load()
rng = np.random.default_rng(3)
timer = dev.Timer()


def timed(fn):
    for _ in range(3):
        fn()
    ms = []
    for _ in range(10):
        timer.start()
        fn()
        timer.stop()
        ms.append(timer.ms())
    return float(np.mean(ms))


print(f"| case | dtype | ms | GB/s (algorithmic) | % of 8 TB/s | bit-exact vs NumPy (first 2^22) |")
print("|---|---|---|---|---|---|")
for dt in ("float32", "float64", "int8", "int16", "int64"):
    s = np.dtype(dt).itemsize
    piece = (rng.normal(0, 100, 1 << 22) if np.dtype(dt).kind == "f" else rng.integers(-100, 100, 1 << 22)).astype(dt)
    a, b, o = dev.empty(n + 64, dt), dev.empty(n + 64, dt), dev.empty(n + 64, dt)
    dp = dev.to_device(piece)
    for x in (a, b):
        for off in range(0, n + 64, piece.size):
            m = min(piece.size, n + 64 - off)
            cabi.check(lib.pncu_memcpy_d2d(x.ptr + off * s, dp.ptr, m * s, None), "fill")
    dev.sync()
    r1 = dev.empty(1, dt)
    ri = dev.empty(1, np.int64)
    k = 1 << 22
    ha = np.concatenate([piece, piece])[: k + 8]
    cases = [
        ("a[1:] - a[:-1]", 3 * s, lambda: dev.binary("SUB", a.view(1, n), a.view(0, n), out=o.view(0, n)),
         lambda: ha[1:k + 1] - ha[:k]),
        ("a[1:] + b[:-1]", 3 * s, lambda: dev.binary("ADD", a.view(1, n), b.view(0, n), out=o.view(0, n)),
         lambda: ha[1:k + 1] + ha[:k]),
        ("a[3:] * b[5:] -> out[1:]", 3 * s, lambda: dev.binary("MUL", a.view(3, n), b.view(5, n), out=o.view(1, n)),
         lambda: ha[3:k + 3] * ha[5:k + 5]),
        ("abs(a[3:])", 2 * s, lambda: dev.unary("ABS", a.view(3, n), out=o.view(0, n)), lambda: np.abs(ha[3:k + 3])),
        ("aligned a + b (reference point)", 3 * s, lambda: dev.binary("ADD", a.view(0, n), b.view(0, n), out=o.view(0, n)),
         lambda: ha[:k] + ha[:k]),
    ]
    if np.dtype(dt).kind == "f":
        cases.append(("sin(a[1:])" if dt == "float32" else "sqrt(abs)(a[1:])", 2 * s,
                      (lambda: dev.unary("SIN", a.view(1, n), out=o.view(0, n))) if dt == "float32" else
                      (lambda: dev.unary("SQRT", a.view(1, n), out=o.view(0, n))), None))
    for name, bpe, fn, want in cases:
        ms = timed(fn)
        gbs = bpe * n / (ms * 1e-3) / 1e9
        ok = "-"
        if want is not None:
            off = 1 if "out[1:]" in name else 0
            got = o.view(off, k).to_host()
            w = want()
            ok = "yes" if got.tobytes() == w.astype(dt).tobytes() else "NO"
        print(f"| {name} | {dt} | {ms:.4f} | {gbs:.0f} | {100 * gbs / 8000:.1f} | {ok} |", flush=True)
    for name, fn, want in (("sum(a[1:])", lambda: dev.reduce("ADD", a.view(1, n), out=r1), None),
                           ("max(a[1:])  (what ndarray.max() calls)", lambda: dev.reduce("MAX", a.view(1, n), out=r1), None),
                           ("argmax(a[1:])", lambda: dev.argminmax(0, a.view(1, n), out=ri), None),
                           ("max(a) aligned (reference point)", lambda: dev.reduce("MAX", a.view(0, n), out=r1), None)):
        ms = timed(fn)
        gbs = s * n / (ms * 1e-3) / 1e9
        tail = lib.pncu_last_tail_ns() / 1e3 if hasattr(lib, "pncu_last_tail_ns") else float("nan")
        print(f"| {name} | {dt} | {ms:.4f} | {gbs:.0f} | {100 * gbs / 8000:.1f} | tail {tail:.1f} us |", flush=True)
    for x in (a, b, o):
        x.free()
