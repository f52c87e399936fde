"""Plain NumPy calls on managed (device-resident) ndarrays through the hooks, spread over all visible GPUs by
the dispatcher: wall clock per call, algorithmic GB/s, the results' bits (so that runs with different GPU
counts can be compared), and the small-call latency with the managed recycler (configs[0]).

    python tools/resident_probe.py [log2_elems ...]         # default 28 30
    PNUMPY_CUDA_DEVICES=1 python tools/resident_probe.py    # the same calls on one GPU
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("PNUMPY_B200_AUTOINIT", "0")
import pnumpy_b200 as pn  # noqa: E402

lgs = [int(x) for x in sys.argv[1:] if not x.startswith('-')] or [28, 30]
pn.init()
pn.enable()
ndev = pn.cuda_info()["devices"]
out = {"gpus": ndev, "sizes": {}}


def med(fn, k=20):
    fn()
    fn()
    ts = []
    for _ in range(k):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts)), float(np.min(ts))


for lg in lgs:
    m = 1 << lg
    ma, mb, mo = (pn.managed_empty(m, np.float32) for _ in range(3))
    piece = np.random.default_rng(1).random(1 << 22, dtype=np.float32) * 253 + 1
    ma.reshape(-1, piece.size)[:] = piece
    mb.reshape(-1, piece.size)[:] = piece[::-1]
    rows = (("add", 12, lambda: np.add(ma, mb, out=mo)), ("sin", 8, lambda: np.sin(ma, out=mo)),
            ("sum", 4, lambda: ma.sum()), ("max", 4, lambda: ma.max()), ("argmax", 4, lambda: ma.argmax()),
            ("muladd_sum", 12, lambda: pn.muladd_sum(ma, mb, mo)))
    res = {}
    print(f"{ndev} GPU(s), 2^{lg} float32 per operand (managed memory)", flush=True)
    for name, bpe, fn in rows:
        s, best = med(fn)
        r = fn()
        bits = np.asarray(r).tobytes().hex() if name not in ("add", "sin") else None
        res[name] = {"ms": round(s * 1e3, 4), "best_ms": round(best * 1e3, 4), "gbs": round(bpe * m / s / 1e9, 1), "bits": bits,
                     "gpus": int(pn.cuda_info()["gpus_last"])}
        print(f"  {name:12s} {s * 1e3:8.4f} ms (best {best * 1e3:8.4f})  {bpe * m / s / 1e9:9.1f} GB/s  gpus={res[name]['gpus']}  {bits or ''}", flush=True)
    # element-wise results: a 64-bit digest computed on the GPU side of things would be self-referential; use the host
    np.add(ma, mb, out=mo)
    if "--trace" in sys.argv or os.environ.get("PROBE_TRACE"):
        from pnumpy_b200 import _pnumpy
        _pnumpy.cuda_trace(True)
        names = ("lock", "plan", "place", "post", "own launch/shard", "join", "result")
        for name, fn in (("add", lambda: np.add(ma, mb, out=mo)), ("sum", lambda: ma.sum()), ("argmax", lambda: ma.argmax())):
            acc, wall = np.zeros(7), []
            for _ in range(20):
                t0 = time.perf_counter_ns()
                fn()
                t1 = time.perf_counter_ns()
                tr = np.array(_pnumpy.cuda_trace(), dtype=np.int64)
                acc += np.diff(tr)[:7]
                wall.append((t1 - t0, tr[0] - t0, t1 - tr[7]))
            w = np.median(np.array(wall), axis=0)
            print(f"  trace {name:7s} wall {w[0] / 1e3:7.1f} us = numpy-in {w[1] / 1e3:5.1f} + " +
                  " + ".join(f"{n_} {v / 20 / 1e3:5.1f}" for n_, v in zip(names, acc)) + f" + numpy-out {w[2] / 1e3:5.1f}", flush=True)
        _pnumpy.cuda_trace(False)
    res["add_crc"] = int(np.bitwise_xor.reduce(np.asarray(mo[: 1 << 24]).view(np.uint32).astype(np.uint64) * np.arange(1, (1 << 24) + 1, dtype=np.uint64)))
    out["sizes"][str(lg)] = res
    del ma, mb, mo

# configs[0]: 1 000 000 float32, a+b, hooks off vs on, arrays born through the managed recycler
pn.recycler_enable(kind="managed")
try:
    n = 1_000_000
    a = (np.arange(n, dtype=np.int64) % 253 + 1).astype(np.float32)
    b = a.copy()
    c = np.empty(n, np.float32)
    small = {}
    for name, fn in (("a+b", lambda: np.add(a, b, out=c)), ("a+5", lambda: np.add(a, np.float32(5), out=c)),
                     ("sin", lambda: np.sin(a, out=c)), ("sum", lambda: a.sum()), ("a==b", lambda: np.equal(a, b)),
                     ("a*b+a (temporaries)", lambda: a * b + a)):
        s, best = med(fn, 200)
        small[name] = {"median_us": round(s * 1e6, 2), "best_us": round(best * 1e6, 2)}
        print(f"  1M managed {name:22s} median {s * 1e6:8.2f} us  best {best * 1e6:8.2f} us", flush=True)
    out["config0_managed"] = small
finally:
    pn.recycler_disable()
pn.disable()
ha = (np.arange(n, dtype=np.int64) % 253 + 1).astype(np.float32)
hb, hc = ha.copy(), np.empty(n, np.float32)
s, best = med(lambda: np.add(ha, hb, out=hc), 200)
out["config0_numpy_us"] = round(s * 1e6, 2)
print(f"  1M NumPy loop a+b median {s * 1e6:8.2f} us", flush=True)
print(json.dumps(out))
