"""Where does the hooked GPU path start to beat NumPy's own loop for HOST operands?  np.add / np.sin / a.sum()
on float32, pageable and pinned, 2^16 .. 2^24 elements, median of 50 calls.  python tools/threshold_probe.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("PNUMPY_B200_AUTOINIT", "0")
import pnumpy_b200 as pn  # noqa: E402


def med(fn, k=50):
    fn()
    ts = []
    for _ in range(k):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts)) * 1e6


pn.init()
pn.enable()
pn.cuda_setthreshold(0)
print("| elements | op | NumPy us | GPU pageable us | GPU pinned us |")
print("|---|---|---|---|---|")
for lg in range(16, 25):
    n = 1 << lg
    a = (np.arange(n, dtype=np.float32) % 253) + 1
    b, o = a.copy(), np.empty(n, np.float32)
    qa, qb, qo = pn.pinned_array(a), pn.pinned_array(a), pn.pinned_empty(n, np.float32)
    for name, f_page, f_pin in (("add", lambda: np.add(a, b, out=o), lambda: np.add(qa, qb, out=qo)),
                                ("sin", lambda: np.sin(a, out=o), lambda: np.sin(qa, out=qo)),
                                ("sum", lambda: a.sum(), lambda: qa.sum())):
        pn.disable()
        t_np = med(f_page)
        pn.enable()
        t_pg, t_pin = med(f_page), med(f_pin)
        print(f"| 2^{lg} | {name} | {t_np:.0f} | {t_pg:.0f} | {t_pin:.0f} |", flush=True)
