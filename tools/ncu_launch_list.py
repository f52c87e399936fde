"""ncu launch-list CSV (--metrics gpu__time_duration.sum --csv) -> markdown table grouped by kernel:
python tools/ncu_launch_list.py gpurun_out/r01_launches.csv > profiles/r01_launches.md (body only)."""
import collections
import sys

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.abspath(__file__)))
from ncu_sweep import read_ncu_csv  # noqa: E402

rows = read_ncu_csv(sys.argv[1])
agg = collections.OrderedDict()
for r in rows:
    name = r["name"].split("(")[0].replace("void ", "").replace("pncu::", "").replace("<unnamed>::", "")
    t = r["metrics"].get("gpu__time_duration.sum", 0.0) * 1e6
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += t
total = sum(a[1] for a in agg.values())
print(f"{len(rows)} launches captured, {total / 1e3:.1f} ms of kernel time in all.\n")
print("| kernel | launches | avg us | total us | share |")
print("|---|---|---|---|---|")
for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{name}` | {c} | {t / c:.1f} | {t:.0f} | {100 * t / total:.1f}% |")
