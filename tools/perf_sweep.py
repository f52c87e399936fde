"""Per-kernel device-resident throughput table for libpnumpy_cuda (CUDA events on the launching
stream, after warm-up).  Usage: python tools/perf_sweep.py [log2_elems] [--ctas 4,8,16] [--json out]
Algorithmic bytes per element follow SURVEY.md 8d."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pnumpy_b200 import cabi, device as dev  # noqa: E402


def time_call(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    t = dev.Timer()
    ms = []
    for _ in range(reps):
        t.start()
        fn()
        t.stop()
        ms.append(t.ms())
    return float(np.mean(ms)), float(np.min(ms))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("lg", nargs="?", type=int, default=28)
    ap.add_argument("--ctas", default="8")
    ap.add_argument("--json", default=None)
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    n = 1 << args.lg
    cabi.init()
    lib = cabi.load()
    peak = 6539.5
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    rows = []
    rng = np.random.default_rng(1234)
    chunk = 1 << 22

    def fill(d, lo, hi):
        host = rng.uniform(lo, hi, chunk).astype(d.dtype) if d.dtype.kind == "f" else rng.integers(1, 254, chunk).astype(d.dtype)
        piece = dev.to_device(host)
        for off in range(0, d.size, chunk):
            m = min(chunk, d.size - off)
            cabi.check(lib.pncu_memcpy_d2d(d.ptr + off * d.dtype.itemsize, piece.ptr, m * d.dtype.itemsize, None), "fill")
        dev.sync()

    for ctas in [int(c) for c in args.ctas.split(",")]:
        lib.pncu_set_tunable(b"ctas_per_sm", ctas)
        lib.pncu_set_tunable(b"reduce_ctas_per_sm", ctas)
        for dt in ("float32", "float64", "int32", "int64", "int8"):
            s = np.dtype(dt).itemsize
            a, b, o = dev.empty(n, dt), dev.empty(n, dt), dev.empty(n, dt)
            ob = dev.empty(n, np.bool_)
            fill(a, 1, 254)
            fill(b, 1, 254)
            cases = [("add", 3 * s, lambda: dev.binary("ADD", a, b, out=o)),
                     ("mul", 3 * s, lambda: dev.binary("MUL", a, b, out=o)),
                     ("greater", 2 * s + 1, lambda: dev.compare("GT", a, b, out=ob)),
                     ("add_scalar", 2 * s, lambda: dev.binary("ADD", a, b.view(5, 1), out=o, b_scalar=True)),
                     ("abs", 2 * s, lambda: dev.unary("ABS", a, out=o)),
                     ("isnan", s + 1, lambda: dev.unary("ISNAN", a, out=ob)),
                     ("sum", s, lambda: dev.reduce("ADD", a)),
                     ("min", s, lambda: dev.reduce("MIN", a)),
                     ("max", s, lambda: dev.reduce("MAX", a)),
                     ("argmax", s, lambda: dev.argminmax(0, a)),
                     ("argmin", s, lambda: dev.argminmax(1, a))]
            if dt in ("float32", "float64"):
                cases += [("divide", 3 * s, lambda: dev.binary("DIV", a, b, out=o)),
                          ("sqrt", 2 * s, lambda: dev.unary("SQRT", a, out=o)),
                          ("sin", 2 * s, lambda: dev.unary("SIN", a, out=o)),
                          ("cos", 2 * s, lambda: dev.unary("COS", a, out=o)),
                          ("log", 2 * s, lambda: dev.unary("LOG", a, out=o)),
                          ("exp", 2 * s, lambda: dev.unary("EXP", b, out=o))]
                fill(b, -80, 80)
            elif dt != "int8":
                of = dev.empty(n, np.float64)
                cases += [("true_divide", 2 * s + 8, lambda: dev.int_true_divide(a, b, out=of))]
            if args.quick:
                cases = [c for c in cases if c[0] in ("add", "greater", "sin", "sum", "argmax", "divide", "true_divide")]
            for name, bpe, fn in cases:
                avg, best = time_call(fn)
                gbs = bpe * n / (avg * 1e-3) / 1e9
                rows.append(dict(op=name, dtype=dt, ctas_per_sm=ctas, ms=avg, ms_best=best, gbs=gbs,
                                 frac_measured=gbs / peak, frac_8tbs=gbs / 8000.0, bytes_per_elem=bpe))
                print(f"ctas={ctas:2d} {name:12s} {dt:8s} {avg:8.4f} ms (best {best:8.4f})  {gbs:8.1f} GB/s  "
                      f"{100 * gbs / peak:5.1f}% of measured {peak:.0f}  {100 * gbs / 8000:5.1f}% of 8 TB/s", flush=True)
            for x in (a, b, o, ob):
                x.free()
    if args.json:
        json.dump(dict(n=n, peak_hbm_gbs=peak, rows=rows), open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
