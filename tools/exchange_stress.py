"""Stress of the cross-GPU exchange inside the one-launch reductions (pnumpy_b200.exchange.Exchange over
pncu_reduce_fused / pncu_argminmax_fused), one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        tools/exchange_stress.py [epochs]

Every epoch each rank sleeps a skewed, pseudo-random time (so that fast ranks run one or two calls ahead of slow
ones: the alternating-table and growing-counter logic), then issues THREE reductions back to back on its stream
without synchronising in between (a different op / window each time), then waits and compares each result with
(i) the plain variant -- pass 1, NCCL all_gather of the block partials, finish kernel -- bit for bit, and (ii) every
other rank's bits.  Any mismatch or exchange timeout exits non-zero.  Prints one JSON line on rank 0.
"""
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("PNUMPY_B200_AUTOINIT", "0")
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from pnumpy_b200 import cabi, device as dev, exchange  # noqa: E402

epochs = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cabi.init()
lib = cabi.load()
cabi.check(lib.pncu_set_device(local), "set_device")
stream = ctypes.c_void_p()
cabi.check(lib.pncu_stream_create(ctypes.byref(stream)), "stream")

ALIGN = 65536
n_max = 64 * ALIGN + 4321          # per-rank shard; windows of it are reduced
rng = np.random.default_rng(1000 + rank)
host = rng.normal(0, 100, n_max + 8 * ALIGN).astype(np.float32)
host[rng.integers(0, host.size, 3)] = np.nan if rank % 3 == 1 else 7.0
buf = dev.to_device(host)
t = cabi.atop_of(np.float32)
xch = exchange.Exchange.from_torch_distributed(max_slots=world * (n_max // 16384 + 2), sync=True)
outs = [dev.empty(1, np.float32) for _ in range(3)] + [dev.empty(1, np.int64) for _ in range(3)]
plain_f, plain_i = dev.empty(1, np.float32), dev.empty(1, np.int64)
delay_rng = np.random.default_rng(77 + 13 * rank)
mismatches, calls = 0, 0
t_start = time.perf_counter()


def plain(kind, fop, shard, counts, lo, target):
    """pass 1, NCCL all_gather of the block partials, finish kernel"""
    pbytes = 16 if kind is not None else lib.pncu_reduce_partial_bytes(fop, t)
    m = max(counts)
    mine = torch.zeros(m * pbytes, dtype=torch.uint8, device="cuda")
    allp = torch.empty(world * m * pbytes, dtype=torch.uint8, device="cuda")
    if kind is None:
        cabi.check(lib.pncu_reduce_partials(fop, t, shard.ptr, shard.size, mine.data_ptr(), stream), "partials")
    else:
        cabi.check(lib.pncu_argminmax_partials(kind, t, shard.ptr, shard.size, ctypes.c_int64(lo), mine.data_ptr(), stream), "arg partials")
    cabi.check(lib.pncu_stream_sync(stream), "sync")
    dist.all_gather_into_tensor(allp, mine)
    packed = torch.cat([allp[r * m * pbytes: r * m * pbytes + counts[r] * pbytes] for r in range(world)])
    torch.cuda.current_stream().synchronize()
    if kind is None:
        cabi.check(lib.pncu_reduce_finish(fop, t, packed.data_ptr(), sum(counts), target.ptr, None, stream), "finish")
    else:
        cabi.check(lib.pncu_argminmax_finish(kind, t, packed.data_ptr(), sum(counts), target.ptr, stream), "arg finish")
    cabi.check(lib.pncu_stream_sync(stream), "sync")
    return target.to_host().tobytes()


for ep in range(epochs):
    # every rank derives the same window sizes from the epoch number; the delays differ per rank
    sizes = [int(((ep * 7919 + r * 104729) % 60 + 3) * ALIGN + ((ep + r) % 5) * 1237) for r in range(world)]
    off = (ep % 8) * ALIGN
    shard = buf.view(off, sizes[rank])
    counts = exchange.block_counts(sizes, np.float32)
    lo = int(sum(sizes[:rank]))
    if delay_rng.random() < 0.5:
        time.sleep(float(delay_rng.random()) * 0.002 * (1 + rank % 3))
    plan = [(("ADD", None), ("MAX", None), (None, 0)), (("MIN", None), (None, 1), ("ADD", None)),
            ((None, 0), (None, 1), ("MAX", None))][ep % 3]
    used = []
    fi, ii = 0, 3
    for op, kind in plan:    # back to back, no synchronisation in between
        if kind is None:
            xch.reduce(op, shard, counts, outs[fi], stream=stream)
            used.append((op, kind, outs[fi]))
            fi += 1
        else:
            xch.argminmax(kind, shard, counts, lo, outs[ii], stream=stream)
            used.append((op, kind, outs[ii]))
            ii += 1
    xch.wait(stream)
    for op, kind, o in used:
        calls += 1
        got = o.to_host().tobytes()
        fop = cabi.BINARY_OPS[op] if op else 0
        want = plain(kind, fop, shard, counts, lo, plain_f if kind is None else plain_i)
        same = got == want
        if kind is None and not same:   # NaN payloads are unspecified
            g, w = np.frombuffer(got, np.float32)[0], np.frombuffer(want, np.float32)[0]
            same = bool(np.isnan(g) and np.isnan(w))
        box = [None] * world
        dist.all_gather_object(box, got)
        agree = all(b == box[0] for b in box) or (kind is None and all(np.isnan(np.frombuffer(b, np.float32)[0]) for b in box))
        if not (same and agree):
            mismatches += 1
            sys.stderr.write(f"rank {rank} epoch {ep} {op or ('argmax', 'argmin')[kind]}: fused {got.hex()} plain {want.hex()} agree={agree}\n")

timeouts = int(lib.pncu_exchange_timeouts())
flag = torch.tensor([mismatches, timeouts], dtype=torch.int64, device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.SUM)
xch.close()
if rank == 0:
    print(json.dumps({"ranks": world, "epochs": epochs, "reductions_checked_per_rank": calls, "mismatches_all_ranks": int(flag[0]),
                      "exchange_timeouts_all_ranks": int(flag[1]), "seconds": round(time.perf_counter() - t_start, 1)}), flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if int(flag[0]) == 0 and int(flag[1]) == 0 else 4)
