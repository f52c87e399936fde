"""Per-kernel HBM traffic under ncu for EVERY covered ufunc of BASELINE.json's configs.

Two halves:
  run    python tools/ncu_sweep.py run [log2_elems] --cases gpurun_out/ncu_cases.json
         launches each (op, dtype) ONCE through the C ABI (no warm-up, no timing: meant to run under
         `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
              --clock-control none --csv --log-file gpurun_out/ncu_sweep.csv`)
         and writes the case list with the number of kernel launches each case made.
  time   python tools/ncu_sweep.py time [log2_elems] --cases gpurun_out/ncu_events.json
         the same cases timed with CUDA events (3 warm-ups, mean of 10), no profiler.
  table  python tools/ncu_sweep.py table gpurun_out/ncu_cases.json gpurun_out/ncu_sweep.csv [events.json]
         joins the ncu CSV with the case list (launch order) and prints a markdown table:
         algorithmic bytes, DRAM bytes ncu counted, ncu duration, algorithmic GB/s, % of 8 TB/s.
         With `events.json` (tools/perf_sweep.py --json, same box, no profiler) the CUDA-event
         numbers are printed beside the ncu ones.
"""
import csv
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def cases_for(dev, np, n, dt, bufs):
    a, b, o, ob, of, c3, r1, ri = bufs
    s = np.dtype(dt).itemsize
    isf = np.dtype(dt).kind == "f"
    c = [("add", 3 * s, lambda: dev.binary("ADD", a, b, out=o)),
         ("subtract", 3 * s, lambda: dev.binary("SUB", a, b, out=o)),
         ("multiply", 3 * s, lambda: dev.binary("MUL", a, b, out=o)),
         ("minimum", 3 * s, lambda: dev.binary("MIN", a, b, out=o)),
         ("maximum", 3 * s, lambda: dev.binary("MAX", a, b, out=o)),
         ("add_scalar", 2 * s, lambda: dev.binary("ADD", a, b.view(5, 1), out=o, b_scalar=True))]
    for name, op in (("equal", "EQ"), ("not_equal", "NE"), ("less", "LT"), ("greater", "GT"),
                     ("less_equal", "LTE"), ("greater_equal", "GTE")):
        c.append((name, 2 * s + 1, lambda op=op: dev.compare(op, a, b, out=ob)))
    c += [("abs", 2 * s, lambda: dev.unary("ABS", a, out=o)),
          ("isnan", s + 1 if isf else 1, lambda: dev.unary("ISNAN", a, out=ob)),
          ("sum", s, lambda: dev.reduce("ADD", a, out=r1)),
          ("min", s, lambda: dev.reduce("MIN", a, out=r1)),
          ("max", s, lambda: dev.reduce("MAX", a, out=r1)),
          ("argmin", s, lambda: dev.argminmax(1, a, out=ri)),
          ("argmax", s, lambda: dev.argminmax(0, a, out=ri)),
          ("muladd", 4 * s, lambda: dev.muladd(a, b, c3, out=o)),
          ("muladd_sum", 3 * s, lambda: dev.muladd_sum(a, b, c3, out=r1))]
    if np.dtype(dt).kind == "f":
        c += [("divide", 3 * s, lambda: dev.binary("DIV", a, b, out=o)),
              ("sqrt", 2 * s, lambda: dev.unary("SQRT", a, out=o)),
              ("sin", 2 * s, lambda: dev.unary("SIN", b, out=o)),
              ("cos", 2 * s, lambda: dev.unary("COS", b, out=o)),
              ("log", 2 * s, lambda: dev.unary("LOG", a, out=o)),
              ("exp", 2 * s, lambda: dev.unary("EXP", b, out=o))]
    else:
        c += [("true_divide", 2 * s + 8, lambda: dev.int_true_divide(a, b, out=of)),
              ("bitwise_and", 3 * s, lambda: dev.binary("BITWISE_AND", a, b, out=o))]
    return c


def run(argv, timed=False):
    import numpy as np
    from pnumpy_b200 import cabi, device as dev
    lg = int(argv[0]) if argv and not argv[0].startswith("-") else 28
    out = argv[argv.index("--cases") + 1] if "--cases" in argv else "gpurun_out/ncu_cases.json"
    n = 1 << lg
    cabi.init()
    lib = cabi.load()
    rng = np.random.default_rng(1234)
    chunk = 1 << 22

    def fill(d, lo, hi):
        host = (rng.uniform(lo, hi, chunk).astype(d.dtype) if d.dtype.kind == "f"
                else (rng.integers(0, 2, chunk) if d.dtype.kind == "b" else rng.integers(1, 120, chunk)).astype(d.dtype))
        piece = dev.to_device(host)
        for off in range(0, d.size, chunk):
            m = min(chunk, d.size - off)
            cabi.check(lib.pncu_memcpy_d2d(d.ptr + off * d.dtype.itemsize, piece.ptr, m * d.dtype.itemsize, None), "fill")
        dev.sync()

    rows = []
    timer = dev.Timer() if timed else None
    dtypes = argv[argv.index("--dtypes") + 1].split(",") if "--dtypes" in argv else ["float32", "float64", "int32", "int64"]
    for dt in dtypes:
        a, b, o = dev.empty(n, dt), dev.empty(n, dt), dev.empty(n, dt)
        ob, of = dev.empty(n, np.bool_), dev.empty(n if np.dtype(dt).kind in "iu" else 16, np.float64)
        c3, r1, ri = dev.empty(n, dt), dev.empty(1, dt), dev.empty(1, np.int64)
        fill(a, 1, 254)
        fill(c3, 1, 254)
        fill(b, -80, 80) if dt.startswith("float") else fill(b, 1, 254)
        if np.dtype(dt).kind in "iu" and np.dtype(dt).itemsize >= 4:
            fill(a, 1, 254)
        for name, bpe, fn in cases_for(dev, np, n, dt, (a, b, o, ob, of, c3, r1, ri)):
            l0 = lib.pncu_launch_count()
            try:
                fn()
            except cabi.PnumpyCudaError:
                continue   # no kernel for this (op, dtype): the reference has none either
            dev.sync()
            row = dict(op=name, dtype=dt, bytes_per_elem=bpe, n=n, launches=int(lib.pncu_launch_count() - l0))
            if timed:  # CUDA events on the launching stream, 3 warm-ups, mean of 10
                for _ in range(3):
                    fn()
                ms = []
                for _ in range(10):
                    timer.start()
                    fn()
                    timer.stop()
                    ms.append(timer.ms())
                row["ms"] = float(np.mean(ms))
                row["gbs"] = bpe * n / (row["ms"] * 1e-3) / 1e9
                print(f"{name:14s} {dt:8s} {row['ms']:8.4f} ms {row['gbs']:8.1f} GB/s {100 * row['gbs'] / 8000:5.1f}% of 8 TB/s", flush=True)
            rows.append(row)
        for x in (a, b, o, ob, of, c3, r1, ri):
            x.free()
    os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
    json.dump(dict(n=n, rows=rows) if timed else rows, open(out, "w"), indent=1)
    print(f"{len(rows)} cases, {sum(r['launches'] for r in rows)} launches -> {out}")


def read_ncu_csv(path):
    """-> list of launches in order: dict(name, metrics{name: value})"""
    lines = open(path, errors="replace").read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    launches = {}
    order = []
    for r in csv.DictReader(lines[start:]):
        i = int(r["ID"])
        if i not in launches:
            launches[i] = dict(name=r["Kernel Name"], metrics={})
            order.append(i)
        v = r["Metric Value"].replace(",", "")
        try:
            v = float(v)
        except ValueError:
            continue
        unit = r.get("Metric Unit", "")
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ns": 1e-9, "us": 1e-6, "usecond": 1e-6,
                 "ms": 1e-3, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}.get(unit, 1.0)
        launches[i]["metrics"][r["Metric Name"]] = v * scale
    return [launches[i] for i in order]


def table(argv):
    cases = json.load(open(argv[0]))
    launches = read_ncu_csv(argv[1])
    events = {}
    if len(argv) > 2:
        for r in json.load(open(argv[2]))["rows"]:
            events[(r["op"], r["dtype"])] = r
    # the sweep's own launches are the LAST sum(launches) of the capture minus nothing: fills are memcpys
    total = sum(c["launches"] for c in cases)
    assert len(launches) >= total, (len(launches), total)
    launches = launches[len(launches) - total:]
    print("| op | dtype | kernel(s) | algorithmic MB | ncu DRAM MB (read + write) | DRAM / algorithmic | ncu time us | "
          "GB/s (algorithmic, ncu time) | % of 8 TB/s | CUDA-event GB/s (no profiler) | % of 8 TB/s |")
    print("|---|---|---|---|---|---|---|---|---|---|---|")
    k = 0
    for c in cases:
        ls = launches[k:k + c["launches"]]
        k += c["launches"]
        alg = c["bytes_per_elem"] * c["n"]
        rd = sum(l["metrics"].get("dram__bytes_read.sum", 0.0) for l in ls)
        wr = sum(l["metrics"].get("dram__bytes_write.sum", 0.0) for l in ls)
        t = sum(l["metrics"].get("gpu__time_duration.sum", 0.0) for l in ls)
        names = " + ".join(l["name"].split("(")[0].replace("<unnamed>::", "").replace("pncu::", "") for l in ls)
        gbs = alg / t / 1e9 if t else 0.0
        ev = events.get((c["op"], c["dtype"]))
        evs = f"{ev['gbs']:.0f} | {100 * ev['gbs'] / 8000:.1f}" if ev else "- | -"
        print(f"| {c['op']} | {c['dtype']} | `{names}` | {alg / 1e6:.0f} | {rd / 1e6:.0f} + {wr / 1e6:.0f} | "
              f"{(rd + wr) / alg:.3f} | {t * 1e6:.1f} | {gbs:.0f} | {100 * gbs / 8000:.1f} | {evs} |")


if __name__ == "__main__":
    if len(sys.argv) < 2 or sys.argv[1] not in ("run", "time", "table"):
        sys.exit(__doc__)
    if sys.argv[1] == "table":
        table(sys.argv[2:])
    else:
        run(sys.argv[2:], timed=sys.argv[1] == "time")
