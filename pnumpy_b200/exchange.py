"""Cross-GPU reductions with the exchange fused into pass 2 (include/pnumpy_cuda.h,
``pncu_reduce_exchange_finish`` / ``pncu_argminmax_exchange_finish``): after the ordinary pass 1 on its own
shard, each participant launches ONE kernel that stores its block partials into every participant's slot
table over NVLink, bumps the arrival counters, waits on the device for everybody else's and folds the
complete table.  No NCCL call, no host synchronisation between the passes; every participant gets the same
bits as a single GPU would produce.

``Exchange`` owns one participant's tables (two, used alternately) and the pointers to everybody else's:
* ``Exchange.in_process(devices, max_slots)`` -- one process driving several GPUs (peer access),
* ``Exchange.from_torch_distributed(max_slots)`` -- one process per GPU: the tables are shared through CUDA
  IPC handles exchanged once with ``torch.distributed.all_gather_object`` (control plane only).
This mirrors what the reference does with its per-chunk partials on the calling thread
(src/pnumpy/_pnumpy.cpp:666-700), spread over GPUs.
"""
import ctypes

import numpy as np

from . import cabi

SLOT_BYTES = 16  # the largest partial: {value bits, index} of an arg-reduction


class Exchange:
    def __init__(self, rank, world, bases, counters, max_slots, device, imported=()):
        self.rank, self.world, self.max_slots, self.device = rank, world, int(max_slots), device
        self._bases, self._counters, self._imported = bases, counters, list(imported)  # [participant][parity] addresses
        self.epoch = 0
        self.expected = [0, 0]
        self._scratch, self._scratch_bytes, self._retired = None, 0, []
        self._tables = []
        for parity in (0, 1):
            t = cabi.PeerTable()
            t.n = world
            for d in range(world):
                t.slots[d] = bases[d][parity]
                t.arrived[d] = counters[d][parity]
            self._tables.append(t)

    # ---- construction ---------------------------------------------------------------------------------
    @staticmethod
    def _alloc(lib, max_slots):
        nbytes = 2 * max_slots * SLOT_BYTES + 256
        p = ctypes.c_void_p()
        cabi.check(lib.pncu_malloc(ctypes.byref(p), nbytes), "pncu_malloc")
        cabi.check(lib.pncu_memset(p, 0, nbytes, None), "pncu_memset")
        cabi.check(lib.pncu_stream_sync(None), "sync")
        return p.value

    @staticmethod
    def _layout(base, max_slots):
        tab = [base + 256, base + 256 + max_slots * SLOT_BYTES]
        return tab, [base, base + 128]

    @classmethod
    def in_process(cls, devices, max_slots):
        """One Exchange per device of `devices`, all in this process (peer access enabled pairwise)."""
        lib = cabi.load()
        cabi.init()
        raw = []
        for d in devices:
            cabi.check(lib.pncu_set_device(d), "set_device")
            for e in devices:
                cabi.check(lib.pncu_enable_peer_access(e), "peer access")
            raw.append(cls._alloc(lib, max_slots))
        lay = [cls._layout(b, max_slots) for b in raw]
        out = []
        for r, d in enumerate(devices):
            x = cls(r, len(devices), [l[0] for l in lay], [l[1] for l in lay], max_slots, d)
            x._own = raw[r]
            out.append(x)
        return out

    @classmethod
    def from_torch_distributed(cls, max_slots, sync=True):
        """This rank's Exchange in a one-process-per-GPU job (torch.distributed initialised, device set).
        With sync=False the caller must synchronise the ranks itself before the first reduction (and can
        turn a failure on one rank -- e.g. CUDA IPC not permitted -- into a collective decision)."""
        import torch.distributed as dist
        lib = cabi.load()
        rank, world = dist.get_rank(), dist.get_world_size()
        dev = ctypes.c_int(0)
        cabi.check(lib.pncu_get_device(ctypes.byref(dev)), "get_device")
        own = cls._alloc(lib, max_slots)
        h = (ctypes.c_ubyte * 64)()
        cabi.check(lib.pncu_ipc_export(own, h), "ipc export")
        box = [None] * world
        dist.all_gather_object(box, bytes(h))
        raw, imported = [], []
        for r in range(world):
            if r == rank:
                raw.append(own)
                continue
            buf = (ctypes.c_ubyte * 64).from_buffer_copy(box[r])
            p = ctypes.c_void_p()
            cabi.check(lib.pncu_ipc_import(buf, ctypes.byref(p)), "ipc import")
            raw.append(p.value)
            imported.append(p.value)
        lay = [cls._layout(b, max_slots) for b in raw]
        x = cls(rank, world, [l[0] for l in lay], [l[1] for l in lay], max_slots, dev.value, imported)
        x._own = own
        if sync:
            dist.barrier()
        return x

    def close(self):
        lib = cabi.load()
        for p in self._imported:
            lib.pncu_ipc_close(p)
        self._imported = []
        lib.pncu_set_device(self.device)
        lib.pncu_device_sync()
        for p in self._retired + ([self._scratch] if self._scratch else []):
            lib.pncu_free(p)
        self._retired, self._scratch, self._scratch_bytes = [], None, 0
        if getattr(self, "_own", None):
            lib.pncu_free(self._own)
            self._own = None

    # ---- one reduction ------------------------------------------------------------------------------------
    def _begin(self, counts, partial_bytes):
        """`counts[r]` = block partials participant r contributes.
        -> (table, slot_base, total, expected, local scratch pointer)"""
        lib = cabi.load()
        parity = self.epoch & 1
        self.epoch += 1
        total = int(sum(counts))
        if total > self.max_slots:
            raise cabi.PnumpyCudaError(f"{total} block partials exceed the exchange tables ({self.max_slots} slots)")
        self.expected[parity] += self.world * lib.pncu_exchange_slices()
        need = int(counts[self.rank]) * partial_bytes
        if getattr(self, "_scratch_bytes", 0) < need:
            p = ctypes.c_void_p()
            cabi.check(lib.pncu_set_device(self.device), "set_device")
            cabi.check(lib.pncu_malloc(ctypes.byref(p), max(need, 1 << 20)), "pncu_malloc")
            if getattr(self, "_scratch", None):
                self._retired.append(self._scratch)   # may still be read by queued work: freed in close()
            self._scratch, self._scratch_bytes = p.value, max(need, 1 << 20)
        return self._tables[parity], int(sum(counts[: self.rank])), total, self.expected[parity], self._scratch

    def reduce(self, op, shard, counts, out, start=None, stream=None):
        """all-reduce of `op` ('ADD', 'MIN', 'MAX', 'MUL') over every participant's `shard` (a DeviceArray)
        into the one-element DeviceArray `out` of this participant.  Asynchronous on `stream`."""
        lib = cabi.load()
        f, t = cabi.BINARY_OPS[op], shard.atype
        table, slot_base, total, expected, scratch = self._begin(counts, lib.pncu_reduce_partial_bytes(f, t))
        cabi.check(lib.pncu_reduce_partials(f, t, shard.ptr, shard.size, scratch, stream), "partials")
        cabi.check(lib.pncu_reduce_exchange_finish(f, t, scratch, int(counts[self.rank]), slot_base, ctypes.byref(table), self.rank,
                                                   total, out.ptr, start.ptr if start is not None else None, expected, stream),
                   "exchange_finish")
        return out

    def argminmax(self, kind, shard, counts, index_base, out, stream=None):
        """global index (int64 DeviceArray `out`) of the first max (kind 0) / min (kind 1); `index_base` is the
        global offset of this participant's shard."""
        lib = cabi.load()
        t = shard.atype
        table, slot_base, total, expected, scratch = self._begin(counts, 16)
        cabi.check(lib.pncu_argminmax_partials(kind, t, shard.ptr, shard.size, int(index_base), scratch, stream), "arg partials")
        cabi.check(lib.pncu_argminmax_exchange_finish(kind, t, scratch, int(counts[self.rank]), slot_base, ctypes.byref(table),
                                                      self.rank, total, out.ptr, expected, stream), "arg exchange_finish")
        return out


def block_counts(shard_sizes, dtype):
    """block partials per participant for shards of `shard_sizes` elements (pncu_reduce_partial_count)"""
    lib = cabi.load()
    t = cabi.atop_of(np.dtype(dtype))
    return [int(lib.pncu_reduce_partial_count(t, int(n))) for n in shard_sizes]
