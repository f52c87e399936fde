"""Per-op end-to-end timing of the hooked path on host ndarrays (pinned and pageable), sweeping the
nte knobs (lanes, slab size).  Usage: python tools/e2e_probe.py [log2_elems]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("PNUMPY_B200_AUTOINIT", "0")
import pnumpy_b200 as pn  # noqa: E402


def best(fn, reps=3):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return min(ts)


def main():
    lg = int(sys.argv[1]) if len(sys.argv) > 1 else 27
    n = 1 << lg
    pn.init()
    pn.enable()
    rng = np.random.default_rng(1)
    piece = rng.uniform(1, 254, 1 << 20)
    for pinned in (True, False):
        alloc = pn.pinned_empty if pinned else (lambda m, dt: np.empty(m, dt))
        arr = {}
        for dt in ("float32", "float64", "int32"):
            a, b, o = alloc(n, dt), alloc(n, dt), alloc(n, dt)
            for off in range(0, n, piece.size):
                a[off:off + piece.size] = piece
                b[off:off + piece.size] = piece[::-1]
            arr[dt] = (a, b, o, alloc(n, np.bool_), alloc(n, np.float64))
        for workers, slab in ((0, 8), (0, 32), (1, 8), (3, 8), (3, 32), (7, 8)):
            if workers == 0:
                pn.thread_disable()
            else:
                pn.thread_enable()
                pn.thread_setworkers(workers)
                for name in ("add", "sin", "greater", "true_divide"):
                    for dt in ("float32", "float64"):
                        pn.atop_setworkers(name, np.dtype(dt).num, workers)
            pn.cuda_setslab(slab << 20)
            for dt in ("float32", "float64"):
                a, b, o, ob, of = arr[dt]
                s = np.dtype(dt).itemsize
                rows = [("add", 3 * s, lambda: np.add(a, b, out=o)),
                        ("greater", 2 * s + 1, lambda: np.greater(a, b, out=ob)),
                        ("sin", 2 * s, lambda: np.sin(a, out=o)),
                        ("sum", s, lambda: a.sum()),
                        ("argmax", s, lambda: a.argmax())]
                for name, bpe, fn in rows:
                    pn.cuda_reset_stats()
                    t = best(fn)
                    info = pn.cuda_info()
                    print(f"{'pinned' if pinned else 'pageable':8s} workers={workers} slab={slab:2d}MiB {name:8s} {dt:8s} "
                          f"{t * 1e3:8.2f} ms  {bpe * n / t / 1e9:7.1f} GB/s algorithmic  lanes={info['lanes_last']} "
                          f"pcie={(info['h2d_bytes'] + info['d2h_bytes']) / max(info['calls'], 1) / t / 1e9:6.1f} GB/s", flush=True)
        # int true_divide goes through NumPy's cast buffers
        a, b, o, ob, of = arr["int32"]
        pn.thread_enable()
        pn.thread_setworkers(3)
        t = best(lambda: np.true_divide(a, b, out=of), reps=2)
        print(f"{'pinned' if pinned else 'pageable':8s} true_divide int32: {t * 1e3:.1f} ms ({16 * n / t / 1e9:.1f} GB/s algorithmic)")
        pn.disable()
        t = best(lambda: np.true_divide(a, b, out=of), reps=2)
        print(f"numpy alone true_divide int32: {t * 1e3:.1f} ms; add f32: {best(lambda: np.add(arr['float32'][0], arr['float32'][1], out=arr['float32'][2])) * 1e3:.1f} ms"
              f"; sin f32: {best(lambda: np.sin(arr['float32'][0], out=arr['float32'][2])) * 1e3:.1f} ms")
        pn.enable()
        del arr


if __name__ == "__main__":
    main()
