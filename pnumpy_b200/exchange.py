"""Cross-GPU reductions with the exchange done INSIDE the reduction kernel (include/pnumpy_cuda.h,
``pncu_reduce_fused`` / ``pncu_argminmax_fused``): each participant launches ONE kernel on its own shard.
Its pass-1 CTAs take tickets; the CTA that draws the last one stores the shard's block partials into every
participant's slot table over NVLink, bumps the arrival counters, waits on the device for everybody else's
and folds the complete table.  No NCCL call, no second launch, no host synchronisation; every participant
gets the same bits as a single GPU would produce.  When every shard starts at a multiple of 64 slots (4 MiB of
input) the owner folds its own 64-slot groups first and only the group partials travel (``hierarchical`` counts
such calls; ``pncu_fused_exchange::group_offset``).

``Exchange`` owns one participant's tables (two, used alternately) and the pointers to everybody else's:
* ``Exchange.in_process(devices, max_slots)`` -- one process driving several GPUs (peer access),
* ``Exchange.from_torch_distributed(max_slots)`` -- one process per GPU: the tables are shared through CUDA
  IPC handles exchanged once with ``torch.distributed.all_gather_object`` (control plane only).
This mirrors what the reference does with its per-chunk partials on the calling thread
(src/pnumpy/_pnumpy.cpp:666-700), spread over GPUs.

A participant that never arrives (a lost rank) is an ERROR, not a hang and not a wrong number: the kernel
gives up after ~2 s WITHOUT folding and without writing the result, flags the call in this participant's
status word, and ``Exchange.wait()`` raises ``PnumpyCudaError``; the Exchange is unusable afterwards (the
arrival counters of the ranks no longer agree).
"""
import ctypes

import numpy as np

from . import cabi

SLOT_BYTES = 16  # the largest partial: {value bits, index} of an arg-reduction
_HEADER = 512    # counters (2 x 128 B) + status word, in front of the two tables


class Exchange:
    def __init__(self, rank, world, bases, counters, max_slots, device, imported=()):
        self.rank, self.world, self.max_slots, self.device = rank, world, int(max_slots), device
        self._bases, self._counters, self._imported = bases, counters, list(imported)  # [participant][parity] addresses
        self.hierarchical = 0    # calls whose shards were group-aligned (group partials travelled, not slots)
        self.epoch = 0
        self.expected = [0, 0]
        self.broken = None
        self._status = None      # device address of this participant's status word
        self._seen_timeouts = None  # pncu_exchange_timeouts() of this device when the first reduction was issued

    # ---- construction ---------------------------------------------------------------------------------
    @staticmethod
    def _group_cap(max_slots):
        return (max_slots + 63) // 64 + 1

    @classmethod
    def _alloc(cls, lib, max_slots):
        nbytes = 2 * (max_slots + cls._group_cap(max_slots)) * SLOT_BYTES + _HEADER
        p = ctypes.c_void_p()
        cabi.check(lib.pncu_malloc(ctypes.byref(p), nbytes), "pncu_malloc")
        cabi.check(lib.pncu_memset(p, 0, nbytes, None), "pncu_memset")
        cabi.check(lib.pncu_stream_sync(None), "sync")
        return p.value

    @classmethod
    def _layout(cls, base, max_slots):
        per = (max_slots + cls._group_cap(max_slots)) * SLOT_BYTES   # one parity: slot table, then group table
        tab = [base + _HEADER, base + _HEADER + per]
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
            x._status = raw[r] + 256
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
        x._status = own + 256
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
        if getattr(self, "_own", None):
            lib.pncu_free(self._own)
            self._own = None

    # ---- one reduction ------------------------------------------------------------------------------------
    def _begin(self, counts):
        """`counts[r]` = block partials participant r contributes -> the pncu_fused_exchange of this call"""
        if self.broken:
            raise cabi.PnumpyCudaError(f"this Exchange is unusable: {self.broken}")
        if self._seen_timeouts is None:
            lib = cabi.load()
            cabi.check(lib.pncu_set_device(self.device), "set_device")
            self._seen_timeouts = int(lib.pncu_exchange_timeouts())
        parity = self.epoch & 1
        self.epoch += 1
        total = int(sum(counts))
        if total > self.max_slots:
            raise cabi.PnumpyCudaError(f"{total} block partials exceed the exchange tables ({self.max_slots} slots)")
        self.expected[parity] += self.world - 1
        x = cabi.FusedExchange()
        x.n, x.self, x.all, x.root = self.world, self.rank, 1, 0
        x.slot_base, x.total_count = int(sum(counts[: self.rank])), total
        # every shard starts at a multiple of the fold group (64 slots = 4 MiB of input): each participant folds its
        # own groups and only the group partials cross NVLink (same bits: pncu_fused_exchange in the header)
        aligned = shards_start_at_whole_groups(counts)
        self.hierarchical += int(aligned)
        for d in range(self.world):
            x.slots[d] = self._bases[d][parity]
            x.arrived[d] = self._counters[d][parity]
        x.group_offset = self.max_slots * SLOT_BYTES if aligned else 0   # _layout: the group table follows the slot table
        x.expected = self.expected[parity]
        # the status word doubles as the completion flag: its TIMEOUT bit is what wait() looks at
        x.done_flag = self._status
        x.done_value = self.epoch
        return x

    def reduce(self, op, shard, counts, out, start=None, stream=None):
        """all-reduce of `op` ('ADD', 'MIN', 'MAX', 'MUL') over every participant's `shard` (a DeviceArray)
        into the one-element DeviceArray `out` of this participant.  Asynchronous on `stream`: call wait()
        before reading `out`."""
        lib = cabi.load()
        f, t = cabi.BINARY_OPS[op], shard.atype
        x = self._begin(counts)
        cabi.check(lib.pncu_reduce_fused(f, t, shard.ptr, shard.size, ctypes.byref(x), out.ptr,
                                         start.ptr if start is not None else None, stream), "reduce_fused")
        return out

    def argminmax(self, kind, shard, counts, index_base, out, stream=None):
        """global index (int64 DeviceArray `out`) of the first max (kind 0) / min (kind 1); `index_base` is the
        global offset of this participant's shard."""
        lib = cabi.load()
        x = self._begin(counts)
        cabi.check(lib.pncu_argminmax_fused(kind, shard.atype, shard.ptr, shard.size, int(index_base), ctypes.byref(x),
                                            out.ptr, stream), "argminmax_fused")
        return out

    def wait(self, stream=None):
        """Synchronise `stream` and raise if any reduction since the last wait() gave up on a participant
        (nothing was written to its `out`).  Returns the number of reductions issued so far."""
        lib = cabi.load()
        cabi.check(lib.pncu_set_device(self.device), "set_device")
        cabi.check(lib.pncu_stream_sync(stream), "sync")
        word = ctypes.c_uint64(0)
        cabi.check(lib.pncu_memcpy_d2h(ctypes.byref(word), self._status, 8, stream), "status")
        cabi.check(lib.pncu_stream_sync(stream), "sync")
        timeouts = int(lib.pncu_exchange_timeouts())   # an earlier call's flag may have been overwritten by a later one
        if word.value & cabi.FUSED_TIMEOUT_BIT or (self._seen_timeouts is not None and timeouts > self._seen_timeouts):
            self.broken = (f"rank {self.rank}: a participant did not deliver its block partials within the bounded wait "
                           f"(reduction #{word.value & ~cabi.FUSED_TIMEOUT_BIT}); no result was produced")
            raise cabi.PnumpyCudaError(self.broken)
        return self.epoch


def shards_start_at_whole_groups(counts):
    """True when every participant's first slot (the running sum of `counts`) is a multiple of the fold group
    (pncu_reduce_group_slots() = 64 slots): level A of pass 2 can then be done by each shard's owner."""
    group = int(cabi.load().pncu_reduce_group_slots())
    base = 0
    for c in counts:
        if base % group:
            return False
        base += int(c)
    return True


def block_counts(shard_sizes, dtype):
    """block partials per participant for shards of `shard_sizes` elements (pncu_reduce_partial_count)"""
    lib = cabi.load()
    t = cabi.atop_of(np.dtype(dtype))
    return [int(lib.pncu_reduce_partial_count(t, int(n))) for n in shard_sizes]
