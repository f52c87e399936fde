#!/usr/bin/env python
"""bench.py -- the atop ufunc hot path on B200, measured the way BASELINE.json asks.

    python bench.py [--gpus N] [--steps K] [--warmup W]        # this repo's CUDA path
    python bench.py --impl reference [...]                     # the reference's threaded CPU path
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   (N > 1)

One STEP = one pass of BASELINE.json configs[1]: the binary ufunc sweep add / multiply / divide /
greater over int32, int64, float32, float64 on 2^28-element device-resident arrays (16 kernel
launches, 272 algorithmic bytes per element).  With N GPUs every rank runs the sweep
on its own 2^28-element shard (weak scaling; element-wise work needs no exchange).
`value` is the whole-job algorithmic GB/s with inputs resident in HBM; `e2e` is the same sweep
through the NumPy ufunc hooks on pinned HOST ndarrays (PCIe copies inside the timed region).
Beside the contract's keys the line carries, at top level:
  `resident`        -- STRONG scaling of the product path: ONE process (rank 0) drives all N GPUs through the NumPy
                       hooks on managed ndarrays (np.add / np.sin / ndarray.sum / argmax at 2^28 and 2^30 elements):
                       the nte dispatcher shards every call over the GPUs; results must equal the one-GPU bits;
  `reductions_2p30` -- configs[3]: sum/min/max/argmin/argmax of 2^30 float32 IN TOTAL with NaNs planted, sharded over
                       the N ranks, one kernel per rank with the NVLink exchange inside (pncu_*_fused), timed beside the
                       NCCL all_gather variant; results must be the one-GPU / NumPy answers;
  `parity`          -- every check above as a boolean.  ANY false (or an exchange timeout) makes the process exit 3,
                       so the driver's multi-GPU runs are parity tests too.
`detail` (outside the timed steps) adds the other covered ufuncs -- sin/log/exp/sqrt (configs[2]), the per-GPU
reductions and the a*b+c -> sum chain (configs[4]) -- each with GB/s and the fraction of both HBM peaks.
Every timing is CUDA events on the launching stream, after warm-up; inputs (>= 1 GiB per operand)
are far larger than the 126 MB L2, so no separate flush is needed.
"""
import argparse
import ctypes
import json
import math
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("PNUMPY_B200_AUTOINIT", "0")

METRIC = "atop_ufunc_sweep_throughput"
UNIT = "GB/s"
N_ELEMS = 1 << 28
SWEEP_DTYPES = ("int32", "int64", "float32", "float64")
SWEEP_OPS = ("add", "multiply", "divide", "greater")
ALIGN = 65536  # pncu_reduce_block_elems(): shard boundaries


def bytes_per_elem(op, dt):
    """Algorithmic bytes per element."""
    s = np.dtype(dt).itemsize
    if op in ("add", "multiply", "subtract"):
        return 3 * s
    if op == "divide":
        return 3 * s if np.dtype(dt).kind == "f" else 2 * s + 8
    if op == "greater":
        return 2 * s + 1
    if op in ("sin", "cos", "log", "exp", "sqrt", "abs"):
        return 2 * s
    if op == "isnan":
        return s + 1
    if op in ("sum", "min", "max", "argmin", "argmax"):
        return s
    raise KeyError(op)


SWEEP_BYTES_PER_ELEM = sum(bytes_per_elem(op, dt) for dt in SWEEP_DTYPES for op in SWEEP_OPS)  # 272


def workload_config(n_gpus):
    return {"workload": "configs[1]: binary ufunc sweep add/multiply/divide/greater x int32/int64/float32/float64, "
                        "2^28 elements per GPU, device-resident",
            "elements_per_gpu": N_ELEMS, "launches_per_step": 16, "algorithmic_bytes_per_element": SWEEP_BYTES_PER_ELEM,
            "l2_policy": "inputs (1-2 GiB per operand) exceed the 126 MB L2; no flush needed",
            "parallelism": f"{n_gpus} x independent shards (no data-path collective)"}


# ---- multi-rank host logic (also exercised on CPU over gloo by tests/test_host_logic.py) ---------------
def shard_bounds(n, rank, world):
    """Contiguous even split of [0, n) at multiples of ALIGN elements."""
    units = -(-n // ALIGN)
    per = -(-units // world)
    return min(n, rank * per * ALIGN), min(n, (rank + 1) * per * ALIGN)


def partial_counts(n, itemsize, world):
    """Block-partial slots each rank produces (one per 64 KiB logical block of its shard)."""
    be = 65536 // itemsize
    out = []
    for r in range(world):
        lo, hi = shard_bounds(n, r, world)
        out.append(-(-(hi - lo) // be) if hi > lo else 0)
    return out


def gather_partials(t, counts):
    """ONE all_gather of every rank's block partials (padded to the longest), compacted into the
    global slot order.  `t`: 1-D tensor of this rank's partials on the backend's device."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return t
    m = max(counts)
    pad = torch.zeros(m, dtype=t.dtype, device=t.device)
    pad[: t.numel()] = t
    allp = torch.empty(world * m, dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(allp, pad)
    return torch.cat([allp[r * m: r * m + counts[r]] for r in range(world)])


def fold_in_order(partials):
    """Fixed-order fold of float64 partials (host stand-in for pncu_reduce_finish)."""
    acc = 0.0
    for v in np.asarray(partials, dtype=np.float64):
        acc = acc + float(v)
    return acc


def combine_argminmax_host(kind, pairs, gather=False):
    """(value, global index) pairs -> NumPy's answer: first NaN wins, else the extreme with the
    smallest index.  kind 0 = argmax, 1 = argmin."""
    if gather:
        import torch.distributed as dist
        if dist.is_initialized() and dist.get_world_size() > 1:
            box = [None] * dist.get_world_size()
            dist.all_gather_object(box, pairs)
            pairs = [p for ps in box for p in ps]
    nans = [p for p in pairs if p[0] != p[0]]
    if nans:
        return min(nans, key=lambda p: p[1])
    best = max(p[0] for p in pairs) if kind == 0 else min(p[0] for p in pairs)
    return min((p for p in pairs if p[0] == best), key=lambda p: p[1])


# ---- helpers ----------------------------------------------------------------------------------------------
def measured_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                       "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---- the CUDA arm -----------------------------------------------------------------------------------------
def run_cuda(args):
    from pnumpy_b200 import cabi, device as dev
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        # a CPU-side group for the final wait: ranks > 0 must not sit in an NCCL barrier (a kernel spinning on
        # their GPU) while rank 0 measures the end-to-end path across all GPUs
        cpu_group = dist.new_group(backend="gloo")
    cabi.init()
    lib = cabi.load()
    cabi.check(lib.pncu_set_device(local), "pncu_set_device")
    stream = ctypes.c_void_p()
    cabi.check(lib.pncu_stream_create(ctypes.byref(stream)), "stream")
    peak, peak_kind = measured_peak()
    n = N_ELEMS
    rng = np.random.default_rng(1234 + rank)

    def filled(dt, lo=1, hi=254):
        """device array of n elements tiled from a 4 Mi-element random piece (reference benchmark
        range arange % 253 + 1, src/pnumpy/benchmark.py:56-66)"""
        d = dev.empty(n, dt)
        piece = 1 << 22
        host = rng.uniform(lo, hi, piece).astype(dt) if np.dtype(dt).kind == "f" else rng.integers(lo, hi, piece).astype(dt)
        p = dev.to_device(host)
        isz = np.dtype(dt).itemsize
        for off in range(0, n, piece):
            cabi.check(lib.pncu_memcpy_d2d(d.ptr + off * isz, p.ptr, min(piece, n - off) * isz, None), "fill")
        dev.sync()
        p.free()
        return d

    # ---- operands of the sweep (about 21 GiB) ---------------------------------------------------------
    ops = []  # (name, dtype, bytes_per_elem, thunk)
    keep = []
    for dt in SWEEP_DTYPES:
        a, b = filled(dt), filled(dt)
        o = dev.empty(n, dt)
        ob = dev.empty(n, np.bool_)
        of = o if np.dtype(dt).kind == "f" else dev.empty(n, np.float64)
        keep += [a, b, o, ob, of]
        isf = np.dtype(dt).kind == "f"
        ops.append(("add", dt, lambda a=a, b=b, o=o: dev.binary("ADD", a, b, out=o, stream=stream)))
        ops.append(("multiply", dt, lambda a=a, b=b, o=o: dev.binary("MUL", a, b, out=o, stream=stream)))
        if isf:
            ops.append(("divide", dt, lambda a=a, b=b, o=o: dev.binary("DIV", a, b, out=o, stream=stream)))
        else:
            ops.append(("divide", dt, lambda a=a, b=b, of=of: dev.int_true_divide(a, b, out=of, stream=stream)))
        ops.append(("greater", dt, lambda a=a, b=b, ob=ob: dev.compare("GT", a, b, out=ob, stream=stream)))

    def barrier():
        cabi.check(lib.pncu_device_sync(), "sync")
        if dist:
            dist.barrier()
            cabi.check(lib.pncu_device_sync(), "sync")

    def step():
        for _, _, fn in ops:
            fn()

    # clocks are sampled from before the warm-up to the end of the timed region (nvidia-smi needs a
    # few hundred ms to start; the warm-up steps are the same load)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        time.sleep(0.5)
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    # per-launch events inside the timed region (cheap; give the per-kernel table and the roofline)
    ev = [[dev.Timer() for _ in ops] for _ in range(args.steps)]
    whole = dev.Timer()
    l0 = lib.pncu_launch_count()
    whole.start(stream)
    for k in range(args.steps):
        for i, (_, _, fn) in enumerate(ops):
            ev[k][i].start(stream)
            fn()
            ev[k][i].stop(stream)
    whole.stop(stream)
    ms_total = whole.ms()
    barrier()
    launches = lib.pncu_launch_count() - l0
    clocks = sampler.stop() if sampler else None
    if dist:
        import torch
        t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = SWEEP_BYTES_PER_ELEM * n * world / (ms_per_step * 1e-3) / 1e9

    per_kernel = []
    for i, (name, dt, _) in enumerate(ops):
        ms = float(np.mean([ev[k][i].ms() for k in range(args.steps)]))
        g = bytes_per_elem(name, dt) * n / (ms * 1e-3) / 1e9
        per_kernel.append({"op": name, "dtype": dt, "ms": round(ms, 4), "gbs": round(g, 1),
                           "frac_of_measured_peak": round(g / peak, 4), "frac_of_8tbs": round(g / 8000.0, 4)})
    # dominant kernel: the 8-byte binary element-wise instantiations (add/multiply/divide on f64, i64)
    dom = [k for k in per_kernel if k["dtype"] in ("int64", "float64") and k["op"] != "greater"]
    dom_ms = float(np.mean([k["ms"] for k in dom]))
    dom_bytes = 24.0 * n
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["ew2_vec_kernel_8byte_bytes_per_launch"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "ew2_vec_kernel<op, 8-byte type, LAY_VV> (add/multiply/divide on float64, int64)",
                "achieved": round(achieved, 1), "peak": peak, "peak_kind": peak_kind, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "frac_of_8tbs": round(achieved / 8000.0, 4),
                "algorithmic_bytes_per_launch": dom_bytes, "share_of_step": round(sum(k["ms"] for k in dom) / ms_per_step, 4),
                "traffic": traffic}

    detail = {}
    if not args.no_detail:
        detail = run_detail(args, dev, cabi, lib, stream, keep, n, peak, rank, world, dist)
    for x in keep:
        x.free()

    # ---- parity gates: every rank checks its own results; any failure anywhere fails the whole job -------------
    parity = {}
    red = detail.pop("reductions_2p30", None) if isinstance(detail, dict) else None
    if red is not None:
        parity.update({"reductions_2p30." + k: bool(v) for k, v in red.get("checks", {}).items()})
    for key in ("chain_fused_device_resident",):
        if key in detail:
            parity[key + ".bit_identical_to_unfused"] = bool(detail[key].get("bit_identical_to_unfused"))
    line = None
    if rank == 0:
        e2e = run_e2e(args) if not args.no_e2e else None
        cpu = run_cpu_baseline(sample_lg=args.cpu_lg) if (world == 1 and not args.no_cpu) else None
        line = {"metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": round(ms_per_step, 4), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "i32/i64/f32/f64 (native width, FMA contraction off)",
                "data": "synthetic", "config": workload_config(world), "clocks": clocks, "gpu_launches": int(launches),
                "frac_of_measured_peak": round(value / world / peak, 4), "frac_of_8tbs": round(value / world / 8000.0, 4),
                "roofline": roofline}
        if e2e is not None:
            resident = e2e.pop("resident", None)
            line["e2e"] = e2e
            if resident is not None:
                line["resident"] = resident
                parity.update({"resident." + k: bool(v) for k, v in resident.get("checks", {}).items()})
            ch = e2e.get("chain") or {}
            if isinstance(ch.get("fused"), dict):
                parity["e2e.chain.fused_bit_identical_to_unfused"] = bool(ch["fused"].get("bit_identical_to_unfused"))
            for k_, v_ in ch.items():
                if k_.startswith("managed_recycler_2p") and isinstance(v_, dict):
                    parity["e2e.chain." + k_ + ".bit_identical_to_fused"] = bool(v_.get("bit_identical_to_fused"))
                    parity["e2e.chain." + k_ + ".no_pcie"] = v_.get("h2d_bytes") == 0
            if isinstance(e2e.get("config0"), dict) and "result_equal" in e2e["config0"]:
                parity["e2e.config0.result_equal"] = bool(e2e["config0"]["result_equal"])
                if "managed_recycler_result_equal" in e2e["config0"]:
                    parity["e2e.config0.managed_recycler_result_equal"] = bool(e2e["config0"]["managed_recycler_result_equal"])
        if red is not None:
            line["reductions_2p30"] = red
        if cpu is not None:
            line["cpu_baseline"] = cpu
        line["parity"] = parity
        line["parity_ok"] = all(parity.values())
        line["per_kernel"] = per_kernel
        line["detail"] = detail
    ok = all(parity.values())
    if dist:
        if rank != 0:
            torch.cuda.synchronize()
        # blocks on the CPU: the other ranks' GPUs stay idle during rank 0's end-to-end / resident phase
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=cpu_group)
        ok = bool(int(flag.item()))
        dist.destroy_process_group()
    if line is not None:
        line["parity_ok"] = bool(ok and line["parity_ok"])
        print(json.dumps(line), flush=True)
    if not ok:
        sys.stderr.write("bench.py: PARITY FAILURE: " + json.dumps({k: v for k, v in parity.items() if not v}) + "\n")
        sys.exit(3)


def run_reductions_2p30(args, dev, cabi, lib, stream, rank, world, dist):
    """configs[3]: sum / min / max / argmin / argmax over 2^30 float32 IN TOTAL, sharded over the `world` ranks at
    multiples of 65536 elements, every shard with one NaN planted for the min/max/arg* runs (SURVEY.md 8d case iv).
    Each rank launches ONE kernel per reduction: pass 1 on its shard, the shard's block partials pushed into every
    rank's slot table over NVLink by the last CTA, which then waits on the device for the other shards and folds the
    whole table (pncu_reduce_fused / pncu_argminmax_fused through pnumpy_b200.exchange.Exchange) -- no NCCL call, no
    host synchronisation.  The plain variant (pass 1, NCCL all_gather of the partials, finish kernel) is timed beside
    it and must give the same bits.  Timing: host wall clock around `reps` calls incl. the exchange, max over ranks."""
    import ctypes
    total = 1 << args.red_lg
    lo, hi = shard_bounds(total, rank, world)
    n = hi - lo
    t = cabi.atop_of(np.float32)
    counts = partial_counts(total, 4, world)
    cnt = counts[rank]
    x = dev.empty(n, np.float32)
    piece = (np.random.default_rng(4321).random(1 << 22, dtype=np.float32) * np.float32(253.0) + np.float32(1.0))
    dp = dev.to_device(piece)
    for off in range(0, n, piece.size):
        cabi.check(lib.pncu_memcpy_d2d(x.ptr + off * 4, dp.ptr, min(piece.size, n - off) * 4, None), "fill")
    dev.sync()
    dp.free()
    res, ri = dev.empty(1, np.float32), dev.empty(1, np.int64)
    reps = 10
    out = {"elements_total": total, "elements_per_rank": n, "ranks": world,
           "timing": f"host wall clock around {reps} calls incl. the exchange, max over ranks"}

    def timed_ms(fn):
        for _ in range(3):
            fn()
        cabi.check(lib.pncu_device_sync(), "sync")
        if dist:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        cabi.check(lib.pncu_device_sync(), "sync")
        ms = (time.perf_counter() - t0) * 1e3 / reps
        if dist:
            import torch
            tt_ = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt_, op=dist.ReduceOp.MAX)
            ms = float(tt_.item())
        return ms

    xch, xch_error = None, None
    if world > 1:
        import torch
        from pnumpy_b200 import exchange
        try:
            xch = exchange.Exchange.from_torch_distributed(max_slots=sum(counts), sync=False)
        except Exception as e:  # e.g. CUDA IPC not permitted in this container: every rank must agree to fall back
            xch_error = str(e)
        okflag = torch.tensor([0 if xch_error else 1], dtype=torch.int32, device="cuda")
        dist.all_reduce(okflag, op=dist.ReduceOp.MIN)
        if int(okflag.item()) == 0:
            if xch is not None:
                xch.close()
            xch, xch_error = None, xch_error or "another rank could not map the peer tables"
        dist.barrier()
    out["exchange"] = ("none (1 GPU): one kernel, the last CTA folds" if world == 1 else
                       ("inside the reduction kernel: the last CTA of every rank stores its shard's block partials into every "
                        "rank's slot table over NVLink (CUDA IPC mappings), waits on the device for the others and folds; "
                        "no NCCL call, no host sync" if xch is not None else
                        f"NCCL all_gather between the passes (fused exchange unavailable: {xch_error})"))

    def two_pass_nccl(kind, fop, target):
        """pass 1, NCCL all_gather of the block partials, finish kernel: the plain way"""
        import torch
        pbytes = 16 if kind is not None else lib.pncu_reduce_partial_bytes(fop, t)
        m = max(counts)
        mine_t = torch.zeros(m * pbytes, dtype=torch.uint8, device="cuda")
        all_t = torch.empty(world * m * pbytes, dtype=torch.uint8, device="cuda")
        packed = torch.empty(sum(counts) * pbytes, dtype=torch.uint8, device="cuda")

        def run():
            if kind is None:
                cabi.check(lib.pncu_reduce_partials(fop, t, x.ptr, n, mine_t.data_ptr(), stream), "partials")
            else:
                cabi.check(lib.pncu_argminmax_partials(kind, t, x.ptr, n, ctypes.c_int64(lo), mine_t.data_ptr(), stream), "arg partials")
            cabi.check(lib.pncu_stream_sync(stream), "sync")
            dist.all_gather_into_tensor(all_t, mine_t)
            off = 0
            for r in range(world):   # compact the padded gather into global slot order
                packed[off: off + counts[r] * pbytes] = all_t[r * m * pbytes: r * m * pbytes + counts[r] * pbytes]
                off += counts[r] * pbytes
            torch.cuda.current_stream().synchronize()
            if kind is None:
                cabi.check(lib.pncu_reduce_finish(fop, t, packed.data_ptr(), sum(counts), target.ptr, None, stream), "finish")
            else:
                cabi.check(lib.pncu_argminmax_finish(kind, t, packed.data_ptr(), sum(counts), target.ptr, stream), "arg finish")
        return run

    def one_launch(kind, name):
        if world == 1 or xch is None:
            if kind is None:
                op = {"sum": "ADD", "min": "MIN", "max": "MAX"}[name]
                return lambda: dev.reduce(op, x, out=res, stream=stream)
            return lambda: dev.argminmax(kind, x, out=ri, stream=stream)
        if kind is None:
            op = {"sum": "ADD", "min": "MIN", "max": "MAX"}[name]
            return lambda: xch.reduce(op, x, counts, res, stream=stream)
        return lambda: xch.argminmax(kind, x, counts, lo, ri, stream=stream)

    # the float64 reference of the sum: the piece repeats, so it is a few float64 sums on the host
    full, rem = divmod(total, piece.size)
    s64 = full * float(np.sum(piece, dtype=np.float64)) + float(np.sum(piece[:rem], dtype=np.float64))
    checks = {}
    for name, kind in (("sum", None), ("min", None), ("max", None), ("argmin", 1), ("argmax", 0)):
        if name == "min":   # from here on: one NaN per shard; rank 0's is the first of the whole job
            nan_at = (1234567 * (rank + 1)) % n
            nanv = np.array([np.nan], dtype=np.float32)
            cabi.check(lib.pncu_memcpy_h2d(x.ptr + 4 * nan_at, nanv.ctypes.data, 4, stream), "plant NaN")
            cabi.check(lib.pncu_stream_sync(stream), "sync")
        fop = cabi.BINARY_OPS[{"sum": "ADD", "min": "MIN", "max": "MAX"}.get(name, "MAX")]
        fn = one_launch(kind, name)
        if world > 1 and xch is None:
            fn = two_pass_nccl(kind, fop, res if kind is None else ri)
        ms = timed_ms(fn)
        if xch is not None:
            xch.wait(stream)
        got = res.to_host() if kind is None else ri.to_host()
        row = {"ms": round(ms, 4), "gbs": round(4.0 * total / (ms * 1e-3) / 1e9, 1), "melem_per_s": round(total / (ms * 1e-3) / 1e6, 0)}
        if name == "sum":
            row["result"] = float(got[0])
            row["within_tolerance"] = bool(abs(float(got[0]) - s64) <= 2.0 ** -23 * abs(s64) + 2.0 ** -50 * s64)
            checks["sum_within_tolerance"] = row["within_tolerance"]
        elif kind is None:
            row["nan_semantics_ok"] = bool(np.isnan(got[0]))
            checks[name + "_nan_ok"] = row["nan_semantics_ok"]
        else:
            row["nan_semantics_ok"] = int(got[0]) == 1234567 % (shard_bounds(total, 0, world)[1])  # global index of rank 0's NaN
            checks[name + "_nan_ok"] = row["nan_semantics_ok"]
        if world > 1 and xch is not None:
            plain_target = dev.empty(1, np.float32 if kind is None else np.int64)
            plain = two_pass_nccl(kind, fop, plain_target)
            row["nccl_all_gather_variant_ms"] = round(timed_ms(plain), 4)
            same = plain_target.to_host().tobytes() == got.tobytes()
            if kind is None and name != "sum":
                same = bool(np.isnan(plain_target.to_host()[0]) and np.isnan(got[0]))   # NaN payloads are unspecified
            row["same_bits_as_nccl_variant"] = bool(same)
            checks[name + "_same_as_nccl"] = bool(same)
        out[name] = row
    if xch is not None:
        out["exchange_timeouts"] = int(lib.pncu_exchange_timeouts())
        checks["no_exchange_timeouts"] = out["exchange_timeouts"] == 0
        xch.close()
    out["checks"] = checks
    x.free()
    return out


def run_detail(args, dev, cabi, lib, stream, keep, n, peak, rank, world, dist):
    """Untimed-by-the-contract extras: configs[2] and configs[3] kernels, device-resident."""
    def timeit(fn, reps=10):
        for _ in range(3):
            fn()
        t = dev.Timer()
        ms = []
        for _ in range(reps):
            t.start(stream)
            fn()
            t.stop(stream)
            ms.append(t.ms())
        return float(np.mean(ms))

    out = {}
    byname = {}
    # keep = [a, b, o, ob, of] per dtype in SWEEP_DTYPES order
    for i, dt in enumerate(SWEEP_DTYPES):
        byname[dt] = keep[5 * i: 5 * i + 5]
    rows = []
    for dt in ("float32", "float64"):
        a, b, o, ob, _ = byname[dt]
        # input ranges of SURVEY.md 8d: sin uniform in [-8192, 8192] (the documented Cephes range),
        # log on (0, 1e6], exp on [-80, 80]; sqrt on the sweep's [1, 254) data.  `b` is refilled in place.
        lo_hi = {"sin": (-8192.0, 8192.0), "log": (1e-6, 1e6), "exp": (-80.0, 80.0), "sqrt": None}
        for name, op in (("sin", "SIN"), ("log", "LOG"), ("exp", "EXP"), ("sqrt", "SQRT")):
            src = a
            if lo_hi[name] is not None:
                piece = np.random.default_rng(5).uniform(lo_hi[name][0], lo_hi[name][1], 1 << 22).astype(dt)
                dp = dev.to_device(piece)
                for off in range(0, n, piece.size):
                    cabi.check(lib.pncu_memcpy_d2d(b.ptr + off * piece.itemsize, dp.ptr, piece.nbytes, None), "fill")
                dev.sync()
                dp.free()
                src = b
            ms = timeit(lambda: dev.unary(op, src, out=o, stream=stream))
            rows.append((name, dt, ms))
        ms = timeit(lambda: dev.unary("ISNAN", a, out=ob, stream=stream))
        rows.append(("isnan", dt, ms))
    for dt in SWEEP_DTYPES:
        a = byname[dt][0]
        r1, ri = dev.empty(1, dt), dev.empty(1, np.int64)
        for name, fn in (("sum", lambda: dev.reduce("ADD", a, out=r1, stream=stream)),
                         ("min", lambda: dev.reduce("MIN", a, out=r1, stream=stream)),
                         ("max", lambda: dev.reduce("MAX", a, out=r1, stream=stream)),
                         ("argmax", lambda: dev.argminmax(0, a, out=ri, stream=stream)),
                         ("argmin", lambda: dev.argminmax(1, a, out=ri, stream=stream))):
            rows.append((name, dt, timeit(fn)))
    out["per_gpu_kernels"] = [
        {"op": name, "dtype": dt, "ms": round(ms, 4), "gbs": round(bytes_per_elem(name, dt) * n / (ms * 1e-3) / 1e9, 1),
         "frac_of_measured_peak": round(bytes_per_elem(name, dt) * n / (ms * 1e-3) / 1e9 / peak, 4),
         "frac_of_8tbs": round(bytes_per_elem(name, dt) * n / (ms * 1e-3) / 1e9 / 8000.0, 4),
         "melem_per_s": round(n / (ms * 1e-3) / 1e6, 0)} for name, dt, ms in rows]

    res = dev.empty(1, np.float32)
    out["reductions_2p30"] = run_reductions_2p30(args, dev, cabi, lib, stream, rank, world, dist)

    # configs[4], device-resident leg: t = a*b; u = t+c; r = u.sum() as three kernels (28 B/element)
    a, b, o = byname["float32"][0], byname["float32"][1], byname["float32"][2]

    def as_f32(x):  # reuse the int32 buffers (same size) as float32 scratch
        v = x.view(0, n)
        v.dtype = np.dtype(np.float32)
        return v
    c, tt = as_f32(byname["int32"][2]), as_f32(byname["int32"][0])

    def chain():
        dev.binary("MUL", a, b, out=tt, stream=stream)
        dev.binary("ADD", tt, c, out=o, stream=stream)
        dev.reduce("ADD", o, out=res, stream=stream)
    ms = timeit(chain)
    out["chain_device_resident"] = {"expr": "t=a*b; u=t+c; u.sum()  (float32, unfused: 3 kernels + finish)", "elements": n,
                                    "ms": round(ms, 4), "gbs": round(28.0 * n / (ms * 1e-3) / 1e9, 1),
                                    "frac_of_8tbs": round(28.0 * n / (ms * 1e-3) / 1e9 / 8000.0, 4)}
    # the same chain fused: one pass, 12 B/element actually moved; bit-identical result
    chain_bits = res.to_host().tobytes()
    ms = timeit(lambda: dev.muladd_sum(a, b, c, out=res, stream=stream))
    out["chain_fused_device_resident"] = {
        "expr": "sum(a*b+c) float32 in one pass (muladd_sum_blocks_kernel + finish)", "elements": n, "ms": round(ms, 4),
        "moved_gbs": round(12.0 * n / (ms * 1e-3) / 1e9, 1), "frac_of_8tbs": round(12.0 * n / (ms * 1e-3) / 1e9 / 8000.0, 4),
        "equivalent_gbs_at_28B_per_elem": round(28.0 * n / (ms * 1e-3) / 1e9, 1),
        "bit_identical_to_unfused": res.to_host().tobytes() == chain_bits}
    ms = timeit(lambda: dev.muladd(a, b, c, out=o, stream=stream))
    out["muladd_f32"] = {"expr": "o=a*b+c float32 in one pass (muladd_vec_kernel)", "elements": n, "ms": round(ms, 4),
                         "gbs": round(16.0 * n / (ms * 1e-3) / 1e9, 1), "frac_of_8tbs": round(16.0 * n / (ms * 1e-3) / 1e9 / 8000.0, 4)}
    return out


def run_resident(pn, args):
    """STRONG scaling of the product path (north_star: "splits large 1-D arrays evenly across up to 8 B200s"): this ONE
    process drives all N GPUs through plain NumPy calls on managed ndarrays -- np.add / np.sin / ndarray.sum / max /
    argmax.  The nte dispatcher keeps shard l of every operand on GPU l and runs one kernel per GPU per call (reductions:
    the block partials meet in GPU 0's table over NVLink inside the same kernels); nothing crosses PCIe.  Wall clock per
    call (median of 20), algorithmic GB/s.  Every result is compared with the SAME call restricted to one GPU
    (thread_disable): the bits must not depend on the number of GPUs."""
    out = {"what": "np.add / np.sin / ndarray.sum / max / argmax through the hooks on managed float32 ndarrays; one process, "
                   "all GPUs; wall clock per call (median of 20)", "gpus": int(pn.cuda_info()["devices"]), "sizes": {}}
    checks = {}

    def med(fn, k=20):
        fn()
        fn()
        ts = []
        for _ in range(k):
            t0 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t0)
        return float(np.median(ts))

    try:
        for lg in args.resident_lg:
            m = 1 << lg
            ma, mb, mo, m1 = (pn.managed_empty(m, np.float32) for _ in range(4))
            piece = np.random.default_rng(7).random(1 << 22, dtype=np.float32) * np.float32(253.0) + np.float32(1.0)
            ma.reshape(-1, piece.size)[:] = piece
            mb.reshape(-1, piece.size)[:] = piece[::-1]
            ma[(3 * m) // 5 + 11] = np.float32(1e6)      # a unique maximum, away from any shard boundary
            rows = {}
            pn.thread_enable()
            pn.cuda_reset_stats()
            for name, bpe, fn in (("add", 12, lambda: np.add(ma, mb, out=mo)), ("sin", 8, lambda: np.sin(ma, out=mo)),
                                  ("sum", 4, lambda: ma.sum()), ("max", 4, lambda: ma.max()), ("argmax", 4, lambda: ma.argmax())):
                s_ = med(fn)
                rows[name] = {"ms": round(s_ * 1e3, 4), "gbs": round(bpe * m / s_ / 1e9, 1), "melem_per_s": round(m / s_ / 1e6, 0),
                              "gpus_used": int(pn.cuda_info()["gpus_last"])}
            rows["h2d_bytes"] = int(pn.cuda_info()["h2d_bytes"])
            # N GPUs vs one GPU: same bits
            np.add(ma, mb, out=mo)
            rN = (ma.sum(), ma.max(), int(ma.argmax()))
            pn.thread_disable()
            np.add(ma, mb, out=m1)
            r1 = (ma.sum(), ma.max(), int(ma.argmax()))
            same_add = bool(np.array_equal(mo, m1))
            np.sin(ma, out=m1)
            pn.thread_enable()
            np.sin(ma, out=mo)
            same_sin = bool(np.array_equal(mo, m1))
            rows["bit_identical_to_one_gpu"] = {"add": same_add, "sin": same_sin, "sum": bool(np.float32(rN[0]).tobytes() == np.float32(r1[0]).tobytes()),
                                                "max": bool(rN[1] == r1[1] == np.float32(1e6)), "argmax": bool(rN[2] == r1[2] == (3 * m) // 5 + 11)}
            for k_, v_ in rows["bit_identical_to_one_gpu"].items():
                checks[f"2p{lg}_{k_}_same_as_one_gpu"] = v_
            checks[f"2p{lg}_no_pcie"] = rows["h2d_bytes"] == 0
            rows["sum_result"] = float(rN[0])
            out["sizes"][f"2p{lg}"] = rows
            del ma, mb, mo, m1
    except Exception as e:  # pragma: no cover
        out["error"] = str(e)
        checks["resident_ran"] = False
    finally:
        pn.thread_enable()
    out["checks"] = checks
    return out


# ---- end to end through the NumPy hooks --------------------------------------------------------------------
def run_e2e(args):
    """The same sweep through the reference-facing API: np.add / np.multiply / np.true_divide /
    np.greater on pinned HOST ndarrays with pnumpy_b200 enabled.  H2D and D2H copies are inside the
    timed region (nte stages 8 MiB slabs through 3-deep rings on separate copy/compute streams)."""
    # the dispatcher spreads a call over the visible GPUs: use exactly the N of `--gpus N`, whatever the box has
    os.environ["PNUMPY_CUDA_DEVICES"] = str(max(1, int(os.environ.get("WORLD_SIZE", args.gpus))))
    import pnumpy_b200 as pn
    try:
        pn.init()
    except ImportError as e:
        return {"error": str(e)}
    pn.enable()
    # one lane per GPU at least; 4 lanes on a single GPU (the reference's main + 3 workers)
    ndev = pn.cuda_info()["devices"]
    extra = max(3, ndev - 1)
    pn.thread_setworkers(extra)
    for name in ("add", "multiply", "true_divide", "greater"):
        for dt in SWEEP_DTYPES:
            pn.atop_setworkers(name, np.dtype(dt).num, extra)
    n = 1 << args.e2e_lg
    rng = np.random.default_rng(99)
    arrs = {}
    for dt in SWEEP_DTYPES:
        a, b = pn.pinned_empty(n, dt), pn.pinned_empty(n, dt)
        piece = rng.uniform(1, 254, 1 << 20).astype(dt)
        for off in range(0, n, piece.size):
            a[off: off + piece.size] = piece[: min(piece.size, n - off)]
            b[off: off + piece.size] = piece[::-1][: min(piece.size, n - off)]
        arrs[dt] = (a, b, pn.pinned_empty(n, dt), pn.pinned_empty(n, np.bool_),
                    pn.pinned_empty(n, np.float64) if np.dtype(dt).kind != "f" else None)

    def sweep():
        for dt in SWEEP_DTYPES:
            a, b, o, ob, of = arrs[dt]
            np.add(a, b, out=o)
            np.multiply(a, b, out=o)
            np.true_divide(a, b, out=o if of is None else of)
            np.greater(a, b, out=ob)

    sweep()  # warm-up: lane resources, slab allocation
    pn.cuda_reset_stats()
    reps = max(1, args.e2e_steps)
    t0 = time.perf_counter()
    for _ in range(reps):
        sweep()
    dt_s = (time.perf_counter() - t0) / reps
    info = pn.cuda_info()
    # configs[4]: a*b+c then sum on pinned float32 host ndarrays (intermediates materialised on the host,
    # as NumPy semantics require): 28 algorithmic B/element, 20 B/element over PCIe H2D, 8 D2H
    chain = None
    try:
        a, b, o = arrs["float32"][0], arrs["float32"][1], arrs["float32"][2]
        c = arrs["int32"][2].view(np.float32)
        t = arrs["int32"][0].view(np.float32)

        def run_chain():
            np.multiply(a, b, out=t)
            np.add(t, c, out=o)
            return o.sum()
        run_chain()
        t0 = time.perf_counter()
        for _ in range(reps):
            r = run_chain()
        ch_s = (time.perf_counter() - t0) / reps
        chain = {"expr": "t=a*b; u=t+c; u.sum() on pinned float32 host ndarrays", "elements": n,
                 "ms": round(ch_s * 1e3, 2), "gbs": round(28.0 * n / ch_s / 1e9, 2), "result": float(r)}
        # the same chain through pn.muladd_sum: 12 B/element over PCIe, no temporaries, same bits
        rf = pn.muladd_sum(a, b, c)
        t0 = time.perf_counter()
        for _ in range(reps):
            rf = pn.muladd_sum(a, b, c)
        fu_s = (time.perf_counter() - t0) / reps
        chain["fused"] = {"expr": "pn.muladd_sum(a, b, c)", "ms": round(fu_s * 1e3, 2),
                          "pcie_gbs": round(12.0 * n / fu_s / 1e9, 2),
                          "equivalent_gbs_at_28B_per_elem": round(28.0 * n / fu_s / 1e9, 2),
                          "bit_identical_to_unfused": bool(np.asarray(rf).tobytes() == np.asarray(r).tobytes())}
    except Exception as e:  # pragma: no cover
        chain = {"error": str(e)}
    # configs[4] at its stated size: 2^30 float32 per operand (4 GiB each; 5 pinned arrays = 20 GiB)
    if args.chain_lg > args.e2e_lg and isinstance(chain, dict) and "error" not in chain:
        try:
            arrs.clear()
            del a, b, o, c, t
            m = 1 << args.chain_lg
            big = [pn.pinned_empty(m, np.float32) for _ in range(5)]
            piece = rng.random(1 << 22, dtype=np.float32) * 253 + 1
            for x in big[:3]:
                x.reshape(-1, piece.size)[:] = piece
            a, b, c, t, o = big

            def run_chain_big():
                np.multiply(a, b, out=t)
                np.add(t, c, out=o)
                return o.sum()
            r = run_chain_big()
            t0 = time.perf_counter()
            for _ in range(reps):
                r = run_chain_big()
            ch_s = (time.perf_counter() - t0) / reps
            rf = pn.muladd_sum(a, b, c)
            t0 = time.perf_counter()
            for _ in range(reps):
                rf = pn.muladd_sum(a, b, c)
            fu_s = (time.perf_counter() - t0) / reps
            chain["at_2p%d" % args.chain_lg] = {
                "elements": m, "unfused_ms": round(ch_s * 1e3, 2), "unfused_gbs": round(28.0 * m / ch_s / 1e9, 2),
                "fused_ms": round(fu_s * 1e3, 2), "fused_pcie_gbs": round(12.0 * m / fu_s / 1e9, 2),
                "fused_equivalent_gbs_at_28B_per_elem": round(28.0 * m / fu_s / 1e9, 2),
                "bit_identical": bool(np.asarray(rf).tobytes() == np.asarray(r).tobytes()), "result": float(r)}
            del big, a, b, c, t, o
        except Exception as e:  # pragma: no cover
            chain["at_2p%d" % args.chain_lg] = {"error": str(e)}
    # configs[4] once more, through the UNMODIFIED NumPy API with the managed recycler on: the operands and the
    # temporaries of `a*b + c` are born in CUDA managed memory (NEP-49 allocator), so after the first touch nothing
    # crosses PCIe -- what the reference's recycler roadmap asks for (src/pnumpy/__init__.py:18-21)
    arrs.clear()
    if isinstance(chain, dict) and "error" not in chain:
        try:
            pn.recycler_enable(kind="managed")
            for lg in sorted({args.e2e_lg, args.chain_lg}):
                m = 1 << lg
                piece = np.random.default_rng(99).random(1 << 22, dtype=np.float32) * 253 + 1
                a, b, c = np.empty(m, np.float32), np.empty(m, np.float32), np.empty(m, np.float32)
                for x in (a, b, c):
                    x.reshape(-1, piece.size)[:] = piece
                h0 = pn.cuda_info()["h2d_bytes"]
                r = (a * b + c).sum()          # first call: pages move to the GPU(s)
                ts = []
                for _ in range(5):
                    t0 = time.perf_counter()
                    r = (a * b + c).sum()
                    ts.append(time.perf_counter() - t0)
                s_ = float(np.median(ts))
                rf = pn.muladd_sum(a, b, c)
                chain["managed_recycler_2p%d" % lg] = {
                    "expr": "(a*b + c).sum() in plain NumPy, arrays and temporaries born managed (recycler_enable(kind='managed'))",
                    "elements": m, "ms": round(s_ * 1e3, 3), "equivalent_gbs_at_28B_per_elem": round(28.0 * m / s_ / 1e9, 1),
                    "h2d_bytes": int(pn.cuda_info()["h2d_bytes"] - h0), "gpus_used": int(pn.cuda_info()["gpus_last"]),
                    "bit_identical_to_fused": bool(np.asarray(rf).tobytes() == np.asarray(r).tobytes()), "result": float(r)}
                del a, b, c
        except Exception as e:  # pragma: no cover
            chain["managed_recycler_error"] = str(e)
        finally:
            pn.recycler_disable()
            pn.recycler_trim()
    # the same hooks on DEVICE-RESIDENT ndarrays (CUDA managed memory), spread over all N GPUs by the dispatcher
    resident = run_resident(pn, args)
    # configs[0]: np.add float32 on 1 000 000 contiguous elements (pn.benchmark's size), hooks off vs on.
    # 12 MB of PCIe traffic per call: latency-bound, the GPU path has to beat ~0.5 ms of NumPy loop.
    config0 = None
    try:
        m = 1_000_000
        def med_us(fn, k=200):
            fn()
            ts = []
            for _ in range(k):
                t0 = time.perf_counter()
                fn()
                ts.append(time.perf_counter() - t0)
            return round(float(np.median(ts)) * 1e6, 1)
        pa = (np.arange(m, dtype=np.float32) % 253) + 1  # the reference's benchmark data (benchmark.py:56-66)
        pb, po = pa.copy(), np.empty(m, np.float32)
        qa, qb, qo = pn.pinned_array(pa), pn.pinned_array(pa), pn.pinned_empty(m, np.float32)
        old_thr = pn.cuda_setthreshold(1 << 16)
        pn.disable()
        t_numpy = med_us(lambda: np.add(pa, pb, out=po))
        pn.enable()
        t_pageable = med_us(lambda: np.add(pa, pb, out=po))
        t_pinned = med_us(lambda: np.add(qa, qb, out=qo))
        same = bool(np.array_equal(qo, pa + pb))
        pn.cuda_setthreshold(old_thr)
        config0 = {"what": "np.add float32, 1 000 000 contiguous elements, median of 200 calls (us)", "numpy_loop_us": t_numpy,
                   "gpu_pageable_operands_us": t_pageable, "gpu_pinned_operands_us": t_pinned,
                   "speedup_pinned_vs_numpy": round(t_numpy / t_pinned, 2), "result_equal": same}
        del qa, qb, qo
        # the same call on arrays NumPy itself allocated while the managed recycler was on (default thresholds)
        pn.recycler_enable(kind="managed")
        try:
            ra = (np.arange(m, dtype=np.float32) % 253) + 1
            rb, ro = ra.copy(), np.empty(m, np.float32)
            np.add(ra, rb, out=ro)
            config0["managed_recycler_us"] = med_us(lambda: np.add(ra, rb, out=ro))
            config0["managed_recycler_temporaries_us"] = med_us(lambda: ra + rb)
            config0["managed_recycler_result_equal"] = bool(np.array_equal(ro, pa + pb))
            config0["speedup_managed_vs_numpy"] = round(t_numpy / config0["managed_recycler_us"], 2)
            del ra, rb, ro
        finally:
            pn.recycler_disable()
            pn.recycler_trim()
    except Exception as e:  # pragma: no cover
        config0 = {"error": str(e)}
    pn.disable()
    return {"value": round(SWEEP_BYTES_PER_ELEM * n / dt_s / 1e9, 2), "unit": UNIT, "elements": n, "steps": reps, "chain": chain, "resident": resident, "config0": config0,
            "ms_per_step": round(dt_s * 1e3, 2), "h2d_bytes_per_step": int(info["h2d_bytes"] // reps),
            "d2h_bytes_per_step": int(info["d2h_bytes"] // reps), "gpu_calls_per_step": int(info["calls"] // reps),
            "staged_bytes_per_step": int(info["staged_bytes"] // reps), "lanes": int(info["lanes_last"]),
            "gpus": int(info["gpus_last"]),
            "api": "numpy ufunc call -> hooked loop -> nte dispatcher (pinned host ndarrays)",
            "note": "bound by the H2D direction of PCIe (2 input operands per op); int32/int64 true_divide runs through the "
                    "added (intN,intN)->float64 loops, operands uncast"}


# ---- the reference's threaded CPU path -----------------------------------------------------------------------
REF_SNIPPET = r'''
import json, os, sys, time
import numpy as np
sys.path.insert(0, sys.argv[1])
lg, reps, maxthreads = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
import pnumpy_ref as pn
pn.initialize()
pn.atop_enable(); pn.thread_enable()
cores = len(os.sched_getaffinity(0))
if maxthreads:
    pn.thread_setworkers(min(15, max(1, cores - 1)))
n = 1 << lg
rng = np.random.default_rng(99)
ops = (("add", np.add), ("multiply", np.multiply), ("true_divide", np.true_divide), ("greater", np.greater))
dts = ("int32", "int64", "float32", "float64")
if maxthreads:
    for name, _ in ops:
        for dt in dts + ("float64",):
            pn.atop_setworkers(name, np.dtype(dt).num, 15)
arrs = {}
for dt in dts:
    a = rng.uniform(1, 254, n).astype(dt); b = a[::-1].copy()
    arrs[dt] = (a, b, np.empty(n, dt), np.empty(n, np.bool_), np.empty(n, np.float64))
def sweep():
    for dt in dts:
        a, b, o, ob, of = arrs[dt]
        np.add(a, b, out=o); np.multiply(a, b, out=o)
        np.true_divide(a, b, out=o if a.dtype.kind == "f" else of)
        np.greater(a, b, out=ob)
sweep()
ts = []
for _ in range(reps):
    t0 = time.perf_counter(); sweep(); ts.append(time.perf_counter() - t0)
a, b, o = arrs["float32"][0], arrs["float32"][1], arrs["float32"][2]
c = b[::-1].copy(); t = np.empty(n, np.float32)
def chain():
    np.multiply(a, b, out=t); np.add(t, c, out=o); return o.sum()
chain()
cs = []
for _ in range(reps):
    t0 = time.perf_counter(); chain(); cs.append(time.perf_counter() - t0)
m = 1000000
pa = (np.arange(m, dtype=np.float32) % 253) + 1; pb = pa.copy(); po = np.empty(m, np.float32)
def med_us(k=200):
    np.add(pa, pb, out=po); v = []
    for _ in range(k):
        t0 = time.perf_counter(); np.add(pa, pb, out=po); v.append(time.perf_counter() - t0)
    return float(np.median(v)) * 1e6
on_us = med_us()
pn.atop_disable(); pn.thread_disable()
off_us = med_us()
print(json.dumps({"config0_on_us": on_us, "config0_off_us": off_us, "seconds": ts, "chain_seconds": cs, "cores": cores, "workers": pn.thread_getworkers(), "n": n, "cpustring": pn.cpustring()}))
'''


def run_cpu_baseline(sample_lg=25, reps=3, maxthreads=1):
    """The reference's own threaded CPU path (its extension built by oracle/build_ref_ext.sh), in a
    subprocess because it patches NumPy's loop tables, on a bounded sample of the same sweep."""
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.exists(os.path.join(ref_dir, "pnumpy_ref", ".built")):
        return {"value": None, "unit": UNIT, "kind": "reference", "cores": 0,
                "sample": "oracle/_ref/pnumpy_ref is not built (make -C oracle)"}
    try:
        r = subprocess.run([sys.executable, "-c", REF_SNIPPET, ref_dir, str(sample_lg), str(reps), str(maxthreads)],
                           capture_output=True, text=True, timeout=600)
        d = json.loads(r.stdout.strip().splitlines()[-1])
    except Exception as e:
        return {"value": None, "unit": UNIT, "kind": "reference", "cores": 0, "sample": f"failed: {e}"}
    best = min(d["seconds"])
    threads = min(16, d["workers"] + 1)
    return {"value": round(SWEEP_BYTES_PER_ELEM * d["n"] / best / 1e9, 2), "unit": UNIT, "cores": threads,
            "host_logical_cpus": d["cores"], "kind": "reference",
            "sample": f"same 16-op sweep on 2^{sample_lg} elements per operand, best of {reps} passes, pnumpy reference "
                      f"extension (oracle/_ref) enabled with thread_setworkers({d['workers']}) and atop_setworkers(.., 15): "
                      f"up to {threads} threads",
            "ms_per_pass": round(best * 1e3, 1), "cpu": d.get("cpustring"),
            "chain_gbs": round(28.0 * d["n"] / min(d["chain_seconds"]) / 1e9, 2) if d.get("chain_seconds") else None,
            "config0": {"what": "np.add float32, 1 000 000 elements, median of 200 calls (us), reference enabled vs disabled",
                        "reference_on_us": round(d["config0_on_us"], 1), "reference_off_us": round(d["config0_off_us"], 1)}
            if "config0_on_us" in d else None}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    cpu = run_cpu_baseline(sample_lg=args.cpu_lg, reps=max(1, min(args.steps, 20)))
    line = {"impl": "reference", "metric": METRIC, "value": cpu["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "i32/i64/f32/f64", "data": "synthetic", "config": workload_config(world),
            "ms_per_step": cpu.get("ms_per_pass"), "cpu_baseline": cpu,
            "e2e": {"value": cpu["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-detail", action="store_true")
    ap.add_argument("--e2e-lg", type=int, default=28, help="log2 elements per operand of the end-to-end sweep")
    ap.add_argument("--chain-lg", type=int, default=30, help="log2 elements of the end-to-end a*b+c -> sum chain (configs[4]: 2^30)")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-lg", type=int, default=28, help="log2 elements per operand of the CPU sample (configs[1]: 2^28)")
    ap.add_argument("--red-lg", type=int, default=30, help="log2 TOTAL elements of the sharded reductions (configs[3]: 2^30)")
    ap.add_argument("--resident-lg", type=int, nargs="*", default=[28, 30], help="log2 elements of the strong-scaling resident leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
