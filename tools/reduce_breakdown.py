"""Where a reduction's time goes: pass 1 alone, pass 2 alone, both (one call), per dtype at 2^28 elements.
    python tools/reduce_breakdown.py [ADD|MIN|MAX] [log2_elems]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("PNUMPY_B200_AUTOINIT", "0")
from pnumpy_b200 import cabi, device as dev
from tools.perf_sweep import time_call
cabi.init(); lib = cabi.load()
OP = sys.argv[1] if len(sys.argv) > 1 else "ADD"
n = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 28)
rng = np.random.default_rng(1)
for dt in ("int8", "int16", "float32", "float64"):
    s = np.dtype(dt).itemsize
    piece = (rng.random(1 << 22) * 100 + 1).astype(dt)
    a = dev.empty(n, dt)
    dp = dev.to_device(piece)
    for off in range(0, n, piece.size):
        lib.pncu_memcpy_d2d(a.ptr + off * s, dp.ptr, piece.nbytes, None)
    dev.sync()
    t, f = cabi.atop_of(dt), cabi.BINARY_OPS[OP]
    pb, nb = lib.pncu_reduce_partial_bytes(f, t), lib.pncu_reduce_partial_count(t, n)
    parts, out = dev.empty(nb * pb, np.uint8), dev.empty(1, dt)
    p1 = time_call(lambda: cabi.check(lib.pncu_reduce_partials(f, t, a.ptr, n, parts.ptr, None), "p1"))
    p2 = time_call(lambda: cabi.check(lib.pncu_reduce_finish(f, t, parts.ptr, nb, out.ptr, None, None), "p2"))
    both = time_call(lambda: dev.reduce(OP, a, out=out))
    ri = dev.empty(1, np.int64)
    ka = time_call(lambda: dev.argminmax(0, a, out=ri))
    ideal = s * n / 7.2e9
    print(f"{dt:8s} partials={nb:6d} ideal@7.2TB/s {ideal:.4f} ms | pass1 {p1[0]:.4f} pass2 {p2[0]:.4f} {OP} {both[0]:.4f} ({s*n/both[0]/1e6:.0f} GB/s) argmax {ka[0]:.4f}", flush=True)
    a.free()
