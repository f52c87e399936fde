"""Does a kernel slow down under sustained load (power/clock), and by how much?
python tools/sustained.py OP DTYPE lo hi [reps]   e.g.  LOG float64 1 254 200
R back-to-back launches on 2^28 device-resident elements, one CUDA event pair per launch, while
nvidia-smi samples SM clock / power / throttle reasons every 20 ms.  Prints the mean of the first
5, the middle and the last 20 launches and the clock/power summary."""
import os
import subprocess
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("PNUMPY_B200_AUTOINIT", "0")
from pnumpy_b200 import cabi, device as dev  # noqa: E402

op, dt, lo, hi = sys.argv[1], sys.argv[2], float(sys.argv[3]), float(sys.argv[4])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 200
n = 1 << 28
cabi.init()
lib = cabi.load()
piece = np.random.default_rng(5).uniform(lo, hi, 1 << 22).astype(dt)
a, o = dev.empty(n, dt), dev.empty(n, dt)
dp = dev.to_device(piece)
for off in range(0, n, piece.size):
    lib.pncu_memcpy_d2d(a.ptr + off * piece.itemsize, dp.ptr, piece.nbytes, None)
dev.sync()
if op in ("ADD", "MUL", "DIV"):
    fn = lambda: dev.binary(op, a, a, out=o)  # noqa: E731
    bpe = 2 * piece.itemsize  # in1 == in2: one read
else:
    fn = lambda: dev.unary(op, a, out=o)  # noqa: E731
    bpe = 2 * piece.itemsize
Q = "clocks.sm,clocks.max.sm,power.draw,power.limit,temperature.gpu,clocks_throttle_reasons.active"
p = subprocess.Popen(["nvidia-smi", f"--query-gpu={Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", "0"],
                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
time.sleep(0.6)
timers = [dev.Timer() for _ in range(reps)]
for t in timers:
    t.start()
    fn()
    t.stop()
ms = np.array([t.ms() for t in timers])
time.sleep(0.1)
p.terminate()
out = p.communicate()[0].splitlines()
rows = [[x.strip() for x in l.split(",")] for l in out if l.count(",") >= 5]
clk = np.array([float(r[0]) for r in rows]) if rows else np.array([0.0])
pw = np.array([float(r[2]) for r in rows if r[2].replace(".", "").isdigit()] or [0.0])
reasons = sorted({r[5] for r in rows})


def g(x):
    return bpe * n / (x * 1e-3) / 1e9


print(f"{op} {dt} [{lo},{hi}] x{reps}: first5 {ms[:5].mean():.4f} ms ({g(ms[:5].mean()):.0f} GB/s)  "
      f"mid20 {ms[reps // 2 - 10:reps // 2 + 10].mean():.4f} ({g(ms[reps // 2 - 10:reps // 2 + 10].mean()):.0f})  "
      f"last20 {ms[-20:].mean():.4f} ({g(ms[-20:].mean()):.0f})  min {ms.min():.4f} ({g(ms.min()):.0f})")
print(f"   sm clock MHz min/median/max {clk.min():.0f}/{np.median(clk):.0f}/{clk.max():.0f} (max {rows[0][1] if rows else '?'})  "
      f"power W median/max {np.median(pw):.0f}/{pw.max():.0f} (limit {rows[0][3] if rows else '?'})  reasons {reasons}")
