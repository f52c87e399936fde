"""pn.benchmark() -- the reference's own ratio table (t_numpy / t_pnumpy at 1 000 000 elements, benchmark.py:120-174)
on this box, with pinned operands (the GPU path DMAs them in place) and at 2^24 elements."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("PNUMPY_B200_AUTOINIT", "0")
import pnumpy_b200 as pn  # noqa: E402

pn.init()
pn.enable()
ct = [np.bool_, np.int8, np.int16, np.int32, np.int64, np.float32, np.float64]
print("## 1 000 000 elements, pinned operands, default thresholds")
pn.benchmark(ctypes=ct, pinned=True, loop_size=30)
print("## 16 777 216 elements, pinned operands")
pn.benchmark(ctypes=ct, pinned=True, sizes=[1 << 24], loop_size=10)
print("## 16 777 216 elements, pageable operands")
pn.benchmark(ctypes=ct, pinned=False, sizes=[1 << 24], loop_size=10)
# the reference's own call, nothing added: pn.benchmark() with the managed recycler on -- the arrays benchmark() builds
# with np.arange(...).astype(...) and the outputs NumPy allocates are born in CUDA managed memory
print("## 1 000 000 elements, recycler_enable(kind='managed'), default thresholds (plain pn.benchmark())")
pn.recycler_enable(kind="managed")
try:
    pn.benchmark(ctypes=ct, loop_size=30)
    print("## 16 777 216 elements, recycler_enable(kind='managed')")
    pn.benchmark(ctypes=ct, sizes=[1 << 24], loop_size=10)
finally:
    pn.recycler_disable()
