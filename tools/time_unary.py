"""Time one unary op of libpnumpy_cuda on 2^28 device-resident elements: python tools/time_unary.py LOG float32 lo hi [input offset in elements]
(an offset that is not a multiple of a 32-byte vector exercises the realigning kernel: np.sin(a[1:]))"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("PNUMPY_B200_AUTOINIT", "0")
from pnumpy_b200 import cabi, device as dev
op, dt, lo, hi = sys.argv[1], sys.argv[2], float(sys.argv[3]), float(sys.argv[4])
off_in = int(sys.argv[5]) if len(sys.argv) > 5 else 0
n = 1 << 28
cabi.init(); lib = cabi.load()
piece = np.random.default_rng(5).uniform(lo, hi, 1 << 22).astype(dt)
a, o = dev.empty(n + 64, dt), dev.empty(n, dt)
dp = dev.to_device(piece)
for off in range(0, n, piece.size):
    lib.pncu_memcpy_d2d(a.ptr + off * piece.itemsize, dp.ptr, piece.nbytes, None)
lib.pncu_memcpy_d2d(a.ptr + n * piece.itemsize, dp.ptr, 64 * piece.itemsize, None)
a = a.view(off_in, n)
dev.sync()
for _ in range(3): dev.unary(op, a, out=o)
t = dev.Timer(); ms = []
for _ in range(10):
    t.start(); dev.unary(op, a, out=o); t.stop(); ms.append(t.ms())
m = float(np.mean(ms))
print(f"{os.environ.get('PNUMPY_CUDA_LIB','default'):40s} {op} {dt}: {m:.4f} ms  {2*piece.itemsize*n/m/1e6:.1f} GB/s")
