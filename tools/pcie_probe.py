"""What can the box's PCIe links carry?  Raw cudaMemcpyAsync between pinned host memory and G GPUs at once
(G = 1, 2, 4, 8 as available): H2D alone, D2H alone, both directions together.  The ceiling the end-to-end
(host ndarray) numbers of bench.py are to be read against.   python tools/pcie_probe.py [MiB per GPU]"""
import ctypes
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("PNUMPY_B200_AUTOINIT", "0")
from pnumpy_b200 import cabi  # noqa: E402

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
nbytes = mib << 20
ndev = cabi.init()
lib = cabi.load()


def vp():
    return ctypes.c_void_p()


host_in, host_out, dev_in, dev_out, s_up, s_dn = [], [], [], [], [], []
for d in range(ndev):
    cabi.check(lib.pncu_set_device(d), "set_device")
    hi, ho, di, do, su, sd = vp(), vp(), vp(), vp(), vp(), vp()
    cabi.check(lib.pncu_malloc_host(ctypes.byref(hi), nbytes), "malloc_host")
    cabi.check(lib.pncu_malloc_host(ctypes.byref(ho), nbytes), "malloc_host")
    ctypes.memset(hi, 1, nbytes)
    ctypes.memset(ho, 0, nbytes)
    cabi.check(lib.pncu_malloc(ctypes.byref(di), nbytes), "malloc")
    cabi.check(lib.pncu_malloc(ctypes.byref(do), nbytes), "malloc")
    cabi.check(lib.pncu_stream_create(ctypes.byref(su)), "stream")
    cabi.check(lib.pncu_stream_create(ctypes.byref(sd)), "stream")
    host_in.append(hi); host_out.append(ho); dev_in.append(di); dev_out.append(do); s_up.append(su); s_dn.append(sd)


def run(g, up, down, reps=4):
    def once():
        for d in range(g):
            lib.pncu_set_device(d)
            if up:
                cabi.check(lib.pncu_memcpy_h2d(dev_in[d], host_in[d], nbytes, s_up[d]), "h2d")
            if down:
                cabi.check(lib.pncu_memcpy_d2h(host_out[d], dev_out[d], nbytes, s_dn[d]), "d2h")
        for d in range(g):
            lib.pncu_set_device(d)
            if up:
                cabi.check(lib.pncu_stream_sync(s_up[d]), "sync")
            if down:
                cabi.check(lib.pncu_stream_sync(s_dn[d]), "sync")
    once()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        once()
        best = min(best, time.perf_counter() - t0)
    return (int(up) + int(down)) * g * nbytes / best / 1e9


print(f"{ndev} GPU(s), {mib} MiB per GPU per direction, pinned host memory, best of 4")
print("| GPUs | H2D GB/s | D2H GB/s | both directions GB/s |")
print("|---|---|---|---|")
g = 1
while g <= ndev:
    print(f"| {g} | {run(g, True, False):.1f} | {run(g, False, True):.1f} | {run(g, True, True):.1f} |", flush=True)
    g *= 2
