"""Sustained throughput of all 20 transcendental ufuncs x {float32, float64} (tools/sustained.py protocol,
60 back-to-back launches, 2^28 elements) -> markdown rows.  python tools/trig_table.py > gpurun_out/trig_table.md"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = [("SIN", -80, 80), ("COS", -80, 80), ("TAN", -80, 80), ("ASIN", -1, 1), ("ACOS", -1, 1), ("ATAN", -80, 80),
         ("SINH", -8, 8), ("COSH", -8, 8), ("TANH", -8, 8), ("ASINH", -80, 80), ("ACOSH", 1, 80), ("ATANH", -0.9, 0.9),
         ("LOG", 0.001, 1e6), ("LOG2", 0.001, 1e6), ("LOG10", 0.001, 1e6), ("EXP", -80, 80), ("EXP2", -80, 80),
         ("EXPM1", -8, 8), ("LOG1P", 0, 254), ("CBRT", 1, 254), ("SQRT", 1, 254)]
print("| ufunc | float32 GB/s (mid / last 20 launches) | % of 8 TB/s | float64 GB/s (mid / last 20) | % of 8 TB/s |")
print("|---|---|---|---|---|")
for op, lo, hi in CASES:
    cells = []
    for dt in ("float32", "float64"):
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sustained.py"), op, dt, str(lo), str(hi), "60"],
                           capture_output=True, text=True, timeout=300)
        m = re.search(r"mid20 [0-9.]+ \((\d+)\)\s+last20 [0-9.]+ \((\d+)\)", r.stdout)
        if not m:
            cells += ["?", "?"]
            continue
        mid, last = int(m.group(1)), int(m.group(2))
        cells += [f"{mid} / {last}", f"{100 * min(mid, last) / 8000:.1f}"]
    print(f"| {op.lower()} | {cells[0]} | {cells[1]} | {cells[2]} | {cells[3]} |", flush=True)
