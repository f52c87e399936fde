"""Time one reduction of libpnumpy_cuda on 2^28 device-resident elements: python tools/time_reduce.py ADD float32"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("PNUMPY_B200_AUTOINIT", "0")
from pnumpy_b200 import cabi, device as dev  # noqa: E402

op, dt = sys.argv[1], sys.argv[2]
n = 1 << 28
cabi.init()
lib = cabi.load()
piece = np.random.default_rng(5).uniform(1, 254, 1 << 22).astype(dt)
a = dev.empty(n, dt)
dp = dev.to_device(piece)
for off in range(0, n, piece.size):
    lib.pncu_memcpy_d2d(a.ptr + off * piece.itemsize, dp.ptr, piece.nbytes, None)
dev.sync()
out = dev.empty(1, dt if op != "ARGMAX" else np.int64)
fn = (lambda: dev.argminmax(0, a, out=out)) if op == "ARGMAX" else (lambda: dev.reduce(op, a, out=out))
for _ in range(3):
    fn()
t = dev.Timer()
ms = []
for _ in range(10):
    t.start(); fn(); t.stop(); ms.append(t.ms())
m = float(np.mean(ms))
print(f"{op} {dt}: {m:.4f} ms  {piece.itemsize * n / m / 1e6:.1f} GB/s")
