"""Cost of pass 2 alone (pncu_reduce_finish / pncu_argminmax_finish: one CTA folds `count` block partials), warm (table in L2)
and cold (a 512 MiB copy in between evicts it, as the streaming pass 1 does): CUDA events around the finish kernel only.
    python tools/finish_probe.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("PNUMPY_B200_AUTOINIT", "0")
from pnumpy_b200 import cabi, device as dev  # noqa: E402

cabi.init()
lib = cabi.load()
big_a, big_b = dev.empty(1 << 29, np.uint8), dev.empty(1 << 29, np.uint8)
timer = dev.Timer()


def timed(fn, cold):
    ms = []
    for i in range(13):
        if cold:
            lib.pncu_memcpy_d2d(big_b.ptr, big_a.ptr, 1 << 29, None)
        timer.start(); fn(); timer.stop()
        if i >= 3:
            ms.append(timer.ms())
    return float(np.median(ms)) * 1e3


for dt, op in (("float32", "ADD"), ("float32", "MAX"), ("float64", "MAX"), ("float32", "ARGMAX"), ("int8", "ADD")):
    t = cabi.atop_of(dt)
    for count in (64, 256, 2048, 16384, 65536):
        if op == "ARGMAX":
            pb = 16
            parts = dev.to_device(np.zeros(count * 2, np.int64))
            out = dev.empty(1, np.int64)
            fn = lambda: cabi.check(lib.pncu_argminmax_finish(0, t, parts.ptr, count, out.ptr, None), "fin")
        else:
            f = cabi.BINARY_OPS[op]
            pb = lib.pncu_reduce_partial_bytes(f, t)
            parts = dev.to_device(np.zeros(count * pb, np.uint8))
            out = dev.empty(1, dt)
            fn = lambda: cabi.check(lib.pncu_reduce_finish(f, t, parts.ptr, count, out.ptr, None, None), "fin")
        print(f"{op:7s} {dt:8s} count {count:6d} ({pb:2d} B each): warm {timed(fn, False):7.1f} us   cold {timed(fn, True):7.1f} us", flush=True)
