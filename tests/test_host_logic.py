"""CPU tests of the host-side logic that needs no GPU: the control surface of the extension
(reference semantics of enable/disable/thread_* -- src/pnumpy/_pnumpy.cpp:1766-1966,
src/atop/threads.h:852-872), loud failure of init() without a device, and the multi-rank
reduction combine protocol of bench.py over gloo (world_size 2)."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

import pnumpy_b200 as pn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _have_gpu():
    return pn.cuda_info()["devices"] > 0 or os.path.exists("/dev/nvidia0")


def test_public_surface_matches_reference():
    """every name the reference exports for this path (src/pnumpy/__init__.py:27-35) exists"""
    for name in ['initialize', 'atop_enable', 'atop_disable', 'atop_isenabled', 'atop_info', 'atop_setworkers',
                 'cpustring', 'thread_enable', 'thread_disable', 'thread_isenabled', 'thread_getworkers',
                 'thread_setworkers', 'thread_zigzag', 'ledger_enable', 'ledger_disable', 'ledger_isenabled',
                 'ledger_info', 'recycler_enable', 'recycler_disable', 'recycler_isenabled', 'recycler_info',
                 'timer_gettsc', 'timer_getutc', 'benchmark', 'benchmark_func', 'init', 'enable', 'disable',
                 'cpu_count_linux', 'argmin', 'argmax']:
        assert callable(getattr(pn, name)), name


def test_flags_and_worker_clamping():
    old = pn.atop_isenabled()
    assert pn.atop_enable() is None and pn.atop_isenabled() is True
    assert pn.atop_disable() is None and pn.atop_isenabled() is False
    pn.atop_enable() if old else pn.atop_disable()
    assert pn.thread_disable() is None and pn.thread_isenabled() is False
    assert pn.thread_enable() is None and pn.thread_isenabled() is True
    prev = pn.thread_getworkers()
    assert pn.thread_setworkers(5) == prev and pn.thread_getworkers() == 5      # returns the previous value
    assert pn.thread_setworkers(0) == 5 and pn.thread_getworkers() == 1          # "at least one worker"
    assert pn.thread_setworkers(1000) == 1 and pn.thread_getworkers() == 15      # hard maximum 15 + caller
    pn.thread_setworkers(prev)
    assert pn.thread_zigzag() is True and pn.thread_zigzag() is False            # toggles 0 <-> 3
    pn.disable()
    assert not pn.atop_isenabled() and not pn.thread_isenabled()
    pn.enable()
    assert pn.atop_isenabled() and pn.thread_isenabled()
    assert pn.ledger_isenabled() is False
    pn.ledger_enable()
    assert pn.ledger_isenabled() is True
    pn.ledger_disable()
    t0, t1 = pn.timer_gettsc(), pn.timer_gettsc()
    assert t1 >= t0 > 0 and pn.timer_getutc() > 1_600_000_000 * 10**9
    logical, physical = pn.cpu_count_linux()
    assert logical >= 1


def test_init_fails_loudly_without_gpu():
    if _have_gpu():
        pytest.skip("a GPU is present")
    with pytest.raises(ImportError, match="requires an NVIDIA B200"):
        pn.init()
    assert pn.atop_info() is False          # nothing was hooked
    assert pn.cpustring() is None
    with pytest.raises(RuntimeError):
        pn.recycler_enable()
    a = np.arange(10.0)
    assert np.all(a + a == 2 * a)            # NumPy is untouched
    # and a plain `import pnumpy_b200` (auto-init on) refuses to import at all
    env = dict(os.environ, PNUMPY_B200_AUTOINIT="1", PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-c", "import pnumpy_b200"], env=env, capture_output=True, text=True)
    assert r.returncode != 0 and "requires an NVIDIA B200" in r.stderr


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pnumpy_b200")):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                for needle in ("import oracle", "from oracle", "atop_oracle", "libatop_oracle", "libatop_ref",
                               "oracle/_ref", "oracle_"):
                    assert needle not in text, f"{f} references the oracle ({needle})"


def test_multirank_combine_protocol_gloo():
    """bench.py --gpus N combines reduction partials with ONE all_gather of the block partials and a
    fixed-order fold; here two gloo ranks run that host logic on CPU with NumPy standing in for the
    pass-1 kernel, and must agree with each other and with the single-rank answer."""
    code = textwrap.dedent('''
        import os, sys
        import numpy as np
        import torch, torch.distributed as dist
        sys.path.insert(0, os.environ["REPO"])
        import bench
        dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
        rank, world = dist.get_rank(), dist.get_world_size()
        n = 3 * 65536 * world + 1234
        full = np.random.default_rng(7).normal(0, 10, n).astype(np.float32)
        full[5 * 65536 + 9] = np.nan
        lo, hi = bench.shard_bounds(n, rank, world)
        assert lo % 65536 == 0
        shard = full[lo:hi]
        # pass 1 stand-in: one float64 partial per 16384-element logical block (65536 B of float32)
        be = 65536 // 4
        parts = np.array([shard[i:i + be].astype(np.float64).sum() for i in range(0, shard.size, be)])
        allp = bench.gather_partials(torch.from_numpy(parts), bench.partial_counts(n, 4, world))
        total = np.float32(bench.fold_in_order(allp.numpy()))
        single = np.float32(bench.fold_in_order(np.array(
            [full[i:i + be].astype(np.float64).sum() for i in range(0, n, be)])))
        assert (np.isnan(total) and np.isnan(single)) or total.tobytes() == single.tobytes(), (total, single)
        # arg-reduction combine: (value, global index) pairs, first NaN wins
        v, i = bench.combine_argminmax_host(0, [(shard.max() if not np.isnan(shard).any() else np.nan,
                                                 int(np.argmax(shard)) + lo)], gather=True)
        assert i == 5 * 65536 + 9, i
        dist.barrier()
        if rank == 0:
            print("OK")
        dist.destroy_process_group()
    ''')
    env = dict(os.environ, REPO=ROOT, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", WORLD_SIZE="2",
               PNUMPY_B200_AUTOINIT="0")
    procs = [subprocess.Popen([sys.executable, "-c", code], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.PIPE, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, e[-2000:]
    assert "OK" in outs[0][0]


def test_numpy_propagates_an_exception_set_inside_a_replaced_legacy_loop(tmp_path):
    """"the L3 stub converts failure into PyErr_SetString(RuntimeError) ... verify on NumPy 2.3.5
    before relying on it".  This is that verification, on whatever NumPy runs the tests: a loop installed with
    PyUFunc_ReplaceLoopBySignature that sets RuntimeError (taking the GIL first, as _pnumpy.cpp does) makes the
    NumPy call raise that very RuntimeError -- for small calls (GIL held) and large ones (GIL released), for
    element-wise calls and reductions -- and the ufunc keeps working afterwards.  tests/host/legacy_loop_error_probe.c
    is a 40-line extension compiled here with gcc; it runs in a subprocess because it patches np.add."""
    import subprocess
    import sys
    import sysconfig
    src = os.path.join(ROOT, "tests", "host", "legacy_loop_error_probe.c")
    so = tmp_path / ("probe" + sysconfig.get_config_var("EXT_SUFFIX"))
    subprocess.check_call(["gcc", "-O1", "-fPIC", "-shared", "-I" + sysconfig.get_paths()["include"], "-I" + np.get_include(),
                           src, "-o", str(so)])
    code = r'''
import sys
sys.path.insert(0, sys.argv[1])
import numpy as np, probe
assert probe.install() == 0
for n in (10, 300000):
    a = np.ones(n, np.float32)
    for call in (lambda: np.add(a, a), lambda: a.sum()):
        probe.setmode(1)
        try:
            call()
            raise SystemExit("no exception for n=%d" % n)
        except RuntimeError as e:
            assert "injected failure" in str(e), e
        probe.setmode(0)
        assert float(np.add(a, a)[0]) == 2.0 and float(a.sum()) == n
print("ok")
'''
    r = subprocess.run([sys.executable, "-c", code, str(tmp_path)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == "ok", (r.stdout, r.stderr)


def test_exchange_group_alignment_rule():
    """pncu_fused_exchange::group_offset may only be set when EVERY shard starts at a whole fold group (64 slots);
    the last shard may be ragged."""
    from pnumpy_b200 import cabi, exchange
    assert cabi.load().pncu_reduce_group_slots() == 64
    assert exchange.shards_start_at_whole_groups([64, 128, 7])
    assert exchange.shards_start_at_whole_groups([320, 320, 320, 1])
    assert exchange.shards_start_at_whole_groups([5])
    assert not exchange.shards_start_at_whole_groups([24, 24])
    assert not exchange.shards_start_at_whole_groups([64, 65, 64])
    # block_counts needs no GPU either: slots per shard = ceil(elements / (64 KiB / itemsize))
    assert exchange.block_counts([1 << 20, (1 << 20) + 1], "float32") == [64, 65]
    assert exchange.block_counts([1 << 20], "float64") == [128]
