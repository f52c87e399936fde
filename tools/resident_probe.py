"""Plain NumPy calls on large managed (device-resident) ndarrays, spread over all visible GPUs by the dispatcher:
wall clock per call and algorithmic GB/s.   python tools/resident_probe.py [log2_elems]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("PNUMPY_B200_AUTOINIT", "0")
import pnumpy_b200 as pn  # noqa: E402

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 30
pn.init()
pn.enable()
m = 1 << lg
ma, mb, mo = (pn.managed_empty(m, np.float32) for _ in range(3))
piece = np.random.default_rng(1).random(1 << 22, dtype=np.float32) * 253 + 1
ma.reshape(-1, piece.size)[:] = piece
mb.reshape(-1, piece.size)[:] = piece[::-1]


def t(fn, k=10):
    fn()
    fn()
    t0 = time.perf_counter()
    for _ in range(k):
        fn()
    return (time.perf_counter() - t0) / k


rows = (("np.add", 12, lambda: np.add(ma, mb, out=mo)), ("np.sin", 8, lambda: np.sin(ma, out=mo)),
        ("np.greater", 9, lambda: np.greater(ma, mb)), ("ndarray.sum", 4, lambda: ma.sum()),
        ("ndarray.argmax", 4, lambda: ma.argmax()), ("pn.muladd_sum", 12, lambda: pn.muladd_sum(ma, mb, mo)))
print(f"{pn.cuda_info()['devices']} GPU(s), 2^{lg} float32 per operand (managed memory)")
for name, bpe, fn in rows:
    if name == "np.greater":
        continue  # its bool output is allocated by NumPy in pageable memory: not a resident call
    s = t(fn)
    print(f"{name:16s} {s * 1e3:8.3f} ms  {bpe * m / s / 1e9:9.1f} GB/s  gpus={pn.cuda_info()['gpus_last']}", flush=True)
