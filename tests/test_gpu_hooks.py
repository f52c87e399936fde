"""GPU tests of the drop-in path: NumPy ufunc call -> hooked loop stub -> nte dispatcher (pinned
staging, lanes) -> libpnumpy_cuda kernels -> result in the caller's ndarray.  The checker is NumPy's
own loop reached through the very same hooks with ``pn.disable()`` (what the reference's
tests/test_ufuncs.py:59-88 does with threading on/off), plus the CPU oracle where the reference
and NumPy differ.  Every test asserts that the GPU path really ran (dispatch counters)."""
import json
import os

import numpy as np
import pytest

from inputs import fill_random
from util import DTYPES, FP_DTYPES, assert_bits_equal, ulp_error

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def pn():
    import pnumpy_b200 as pn
    pn.init()  # raises ImportError without a B200
    pn.enable()
    old = pn.cuda_setthreshold(0)  # every call goes to the GPU, however small
    yield pn
    pn.cuda_setthreshold(old)
    pn.disable()  # later test modules use NumPy as a reference: leave its own loops in charge


def _both(pn, fn):
    """(GPU result, NumPy-loop result, GPU calls made)"""
    pn.enable()
    c0 = pn.cuda_info()["calls"]
    with np.errstate(all="ignore"):
        got = fn()
    calls = pn.cuda_info()["calls"] - c0
    pn.disable()
    try:
        with np.errstate(all="ignore"):
            want = fn()
    finally:
        pn.enable()
    return got, want, calls


def test_hook_table_matches_reference(pn):
    """atop_info() lists exactly the (ufunc, dtype) pairs the reference hooks (359 on NumPy 2.3)."""
    want = {(k, d) for k, d in json.load(open(os.path.join(ROOT, "tests", "golden", "hook_keys.json")))}
    got = set(pn.atop_info().keys())
    assert got == want, (sorted(got - want)[:5], sorted(want - got)[:5])
    info = pn.atop_info("add", np.dtype(np.float32).num)
    assert set(info) == {"MaxThreads", "MinElementsToThread", "pBinaryFunc", "pOldFunc", "pReduceFunc", "pUnaryFunc"}
    assert info["MaxThreads"] == 3 and info["pBinaryFunc"] and info["pOldFunc"] and info["pReduceFunc"]
    assert pn.atop_info("sin", np.dtype(np.float64).num)["MaxThreads"] == 7
    assert pn.atop_info("nosuchufunc", 11) is False
    assert "B200" in pn.cpustring()


def test_enable_disable_roundtrip(pn):
    """reference: tests/test_pnumpy.py:10-25"""
    old = pn.atop_isenabled()
    pn.atop_enable()
    assert pn.atop_isenabled() is True
    pn.atop_disable()
    assert pn.atop_isenabled() is False
    a = np.ones(1 << 16, np.float32)
    c0 = pn.cuda_info()["calls"]
    assert np.all(a + a == 2)  # hooks stay installed, the NumPy loop runs
    assert pn.cuda_info()["calls"] == c0
    pn.atop_enable() if old else pn.atop_disable()
    assert pn.atop_isenabled() == old
    prev = pn.atop_setworkers("add", np.dtype(np.float32).num, 5)
    assert prev == 3 and pn.atop_setworkers("add", np.dtype(np.float32).num, prev) == 5
    assert pn.atop_setworkers("add", np.dtype(np.float32).num, 64) is False


def test_small_result(pn, rng):
    """reference: tests/test_pnumpy.py:28-48"""
    m = rng.integers(100, size=(10, 10), dtype=np.int32)
    o = np.empty_like(m)
    for i in range(m.shape[0]):
        for j in range(m.shape[1]):
            o[i, j] = m[i, j] + m[i, j]
    assert np.all(np.add(m, m) == o)


BIN = ["add", "subtract", "multiply", "true_divide", "floor_divide", "minimum", "maximum", "bitwise_and",
       "bitwise_or", "bitwise_xor", "left_shift", "right_shift", "logical_and", "logical_or"]
CMP = ["equal", "not_equal", "greater", "greater_equal", "less", "less_equal"]
UNA = ["abs", "negative", "sign", "floor", "ceil", "trunc", "rint", "sqrt", "isnan", "isinf", "isfinite",
       "logical_not", "invert", "signbit"]


@pytest.mark.parametrize("dt", DTYPES)
def test_ufuncs_gpu_vs_numpy_loops(pn, rng, dt):
    """The reference's sweep (1024x1024 random bit patterns, tests/test_ufuncs.py:59-88) restricted to
    the hooked ufuncs, comparing the GPU loop with NumPy's loop bit for bit."""
    shape = (1024, 1024)
    a = fill_random(shape[0] * shape[1], dt, rng).reshape(shape)
    b = fill_random(shape[0] * shape[1], dt, rng).reshape(shape)
    ran = 0
    for name in BIN + CMP:
        uf = getattr(np, name)
        sig = np.dtype(dt).char * 2
        if not any(t.startswith(sig + "->") for t in uf.types):
            continue
        if (name, np.dtype(dt).num) not in pn.atop_info():
            continue
        bb = b
        if "shift" in name:
            bb = (b.astype(np.int64) % 70 - 3).astype(dt)
        for rhs in (bb, bb[3, 5]):
            got, want, calls = _both(pn, lambda: uf(a, rhs))
            assert_bits_equal(got, want, f"{name} {dt}")
            ran += calls
    for name in UNA:
        uf = getattr(np, name)
        if not any(t.startswith(np.dtype(dt).char + "->") for t in uf.types):
            continue
        if (name, np.dtype(dt).num) not in pn.atop_info():
            continue
        got, want, calls = _both(pn, lambda: uf(a))
        assert_bits_equal(got, want, f"{name} {dt}")
        ran += calls
    assert ran >= 10, "the GPU path did not run"


@pytest.mark.parametrize("dt", FP_DTYPES)
def test_transcendentals_through_hooks(pn, rng, dt):
    n = 1 << 20
    x = {"sin": rng.uniform(-8192, 8192, n), "cos": rng.uniform(-3.14, 3.14, n),
         "log": np.exp(rng.uniform(-80, 80, n)), "exp": rng.uniform(-80, 80, n)}
    from oracle import oracle as O
    for name, v in x.items():
        v = v.astype(dt)
        got, want, calls = _both(pn, lambda: getattr(np, name)(v))
        assert calls >= 1
        if dt == "float32":
            assert_bits_equal(got, O.unary(name.upper(), v), name)  # == the reference's vector body
            if name != "sin":  # sin near zeros inside +-8192: absolute, not relative, accuracy (App. C.2)
                assert ulp_error(got, getattr(np, name)(v.astype(np.float64))).max() <= 2.0
        else:
            truth = getattr(np, name)(v.astype(np.longdouble))
            sp = np.spacing(np.abs(truth).astype(np.float64))
            assert (np.abs(got.astype(np.longdouble) - truth) / sp).max() <= 2.0


@pytest.mark.parametrize("dt", DTYPES)
def test_reductions_through_hooks(pn, rng, dt):
    n = (1 << 20) + 77
    a = fill_random(n, dt, rng)
    if dt in FP_DTYPES:
        a = rng.normal(0, 100, n).astype(dt)
    fns = {"min": lambda: a.min(), "max": lambda: a.max(), "argmax": lambda: a.argmax(), "argmin": lambda: a.argmin()}
    if dt == "bool":
        fns.update({"any": lambda: a.any(), "all": lambda: a.all()})
    else:
        fns["sum"] = lambda: np.add.reduce(a, dtype=a.dtype)
    for name, fn in fns.items():
        got, want, calls = _both(pn, fn)
        assert calls >= 1, f"GPU reduction did not run for {name} {dt}"
        if dt in FP_DTYPES and np.asarray(got).dtype.kind == "f":
            tol = 2.0 ** -22 * abs(float(want)) + 2.0 ** -45 * float(np.abs(a.astype(np.float64)).sum())
            assert abs(float(got) - float(want)) <= tol, (got, want)
        else:
            assert got == want, (got, want)
    if dt in FP_DTYPES:
        for pos in ([0], [n - 1], [70000, 500000]):
            b = a.copy()
            b[pos] = np.nan
            for fn in (lambda: b.min(), lambda: b.max(), lambda: b.argmax(), lambda: b.argmin()):
                got, want, calls = _both(pn, fn)
                assert calls >= 1
                assert (np.isnan(got) and np.isnan(want)) or got == want, (pos, got, want)
        # float32 sum: GPU accumulates in float64 (the reference's intent), within SURVEY 8c tolerance
        got, _, _ = _both(pn, lambda: a.sum())
        ref = float(np.sum(a.astype(np.float64)))
        eps = 2.0 ** -23 if dt == "float32" else 2.0 ** -52
        assert abs(float(got) - ref) <= eps * abs(ref) + 2.0 ** -45 * float(np.abs(a.astype(np.float64)).sum())


def test_layouts_inplace_and_declines(pn, rng):
    n = 1 << 18
    a = rng.normal(size=n).astype(np.float32)
    b = rng.normal(size=n).astype(np.float32)
    want = a + b
    c = a.copy()
    c0 = pn.cuda_info()["calls"]
    c += b                       # in place: out aliases in1
    assert pn.cuda_info()["calls"] == c0 + 1
    assert_bits_equal(c, want, "in-place")
    out = np.empty_like(a)
    np.add(a, b, out=out)
    assert_bits_equal(out, want, "out=")
    # non-contiguous host operands are gathered / scattered by the lanes and run on the GPU
    # (the reference threads its generic strided loop there, src/atop/ops_binary.cpp:563-569)
    c0 = pn.cuda_info()["calls"]
    s = a[::2] + b[::2]
    assert pn.cuda_info()["calls"] == c0 + 1
    assert_bits_equal(s, want[::2].copy(), "strided")
    # partial overlap: NumPy copies the overlapping input itself before calling the loop (and nte
    # would decline a loop-level overlap); either way the answer is NumPy's
    ov = a.copy()
    np.add(ov[:-1], ov[1:], out=ov[1:])
    ref = a.copy()
    pn.disable(); np.add(ref[:-1], ref[1:], out=ref[1:]); pn.enable()
    assert_bits_equal(ov, ref, "overlap")
    # 2-D reduction along axis 1: many short reduce calls, all correct
    m = rng.normal(size=(512, 2048)).astype(np.float64)
    got, wantm, _ = _both(pn, lambda: m.max(axis=1))
    assert_bits_equal(got, wantm, "axis reduce")
    # below the size threshold the old loop runs
    old = pn.cuda_setthreshold(1 << 30)
    c1 = pn.cuda_info()["calls"]
    assert_bits_equal(a + b, want, "below threshold")
    assert pn.cuda_info()["calls"] == c1
    pn.cuda_setthreshold(old)


@pytest.mark.parametrize("dt", ["int16", "int32", "float32", "float64"])
def test_strided_host_operands(pn, rng, dt):
    """Every stride NumPy can hand an inner loop (steps a multiple of the item size, negative included):
    inputs, outputs, in place, reductions -- same bits as NumPy's own loops."""
    n = (1 << 19) + 77
    base_a = (rng.normal(0, 100, 3 * n)).astype(dt)
    base_b = (rng.normal(0, 100, 2 * n) + 1).astype(dt)
    a, b = base_a[::3], base_b[::-2]
    cases = {
        "add": lambda: np.add(a, b),
        "sub_out_strided": lambda: _strided_out(np.subtract, a, b, n, dt),
        "greater": lambda: np.greater(a, b),
        "abs": lambda: np.abs(a),
        "mul_scalar": lambda: a * np.dtype(dt).type(3),
        "neg_reversed": lambda: np.negative(a[::-1]),
        "max": lambda: a.max(),
        "min_rev": lambda: b.min(),
    }
    if dt.startswith("float"):
        cases["sqrt"] = lambda: np.sqrt(np.abs(a))
        cases["sin"] = lambda: np.sin(b)
    for name, fn in cases.items():
        c0 = pn.cuda_info()["calls"]
        got, want, _ = _both(pn, fn)
        assert pn.cuda_info()["calls"] > c0, name
        if name == "sin":  # not NumPy's bits (Cephes / own float64 body): the contiguous GPU result is the reference
            assert_bits_equal(got, np.sin(np.ascontiguousarray(b)), f"sin {dt}")
        else:
            assert_bits_equal(np.asarray(got), np.asarray(want), f"{name} {dt}")
    # in place on a strided view
    v = base_a.copy()
    w = base_a.copy()
    v[::3] += b
    pn.disable(); w[::3] += b; pn.enable()
    assert_bits_equal(v, w, "in-place strided")
    # a column of a C-ordered matrix (stride = row length)
    m = rng.normal(size=(1 << 17, 8)).astype(dt)
    got, want, _ = _both(pn, lambda: m[:, 3] + m[:, 5])
    assert_bits_equal(got, want, "columns")


def _strided_out(ufunc, a, b, n, dt):
    out = np.zeros(2 * n, dtype=dt)
    ufunc(a, b, out=out[1::2])
    return out


def test_lanes_do_not_change_results(pn, rng):
    """thread on/off and worker counts (the reference's tests/test_ufuncs.py protocol): the number of
    lanes must not change a single bit -- reductions included."""
    n = (1 << 22) + 4099
    a = rng.normal(0, 100, n).astype(np.float32)
    b = rng.normal(0, 100, n).astype(np.float32)
    results = []
    oldw = pn.thread_getworkers()
    for threading, workers in ((False, 1), (True, 1), (True, 3), (True, 7)):
        pn.thread_enable() if threading else pn.thread_disable()
        pn.thread_setworkers(workers)
        pn.atop_setworkers("sin", np.dtype(np.float32).num, 7)
        r = (np.add(a, b), np.sin(a), np.greater(a, b), a.sum(), a.min(), a.argmax(), np.float64(a.astype(np.float64).sum()))
        info = pn.cuda_info()
        if not threading:
            assert info["lanes_last"] == 1
        results.append(r)
    pn.thread_enable()
    pn.thread_setworkers(oldw)
    for r in results[1:]:
        for x, y in zip(results[0], r):
            assert np.asarray(x).tobytes() == np.asarray(y).tobytes()
    assert pn.thread_isenabled()


def test_pinned_operands_are_not_staged(pn, rng):
    n = 1 << 22
    a, b, c = pn.pinned_empty(n, np.float32), pn.pinned_empty(n, np.float32), pn.pinned_empty(n, np.float32)
    a[:] = rng.normal(size=n)
    b[:] = rng.normal(size=n)
    pn.cuda_reset_stats()
    np.multiply(a, b, out=c)
    info = pn.cuda_info()
    assert info["calls"] == 1 and info["staged_bytes"] == 0
    assert info["h2d_bytes"] == 2 * 4 * n and info["d2h_bytes"] == 4 * n
    pn.disable()
    want = a * b
    pn.enable()
    assert_bits_equal(np.asarray(c), want, "pinned")
    s = float(c.sum())
    assert abs(s - float(np.sum(want.astype(np.float64)))) <= 1e-6 * n


def test_managed_arrays_run_in_place(pn, rng):
    """Device-resident ndarrays (CUDA managed memory): a ufunc chain over them runs in place on the GPU,
    with no PCIe copies issued by the dispatcher, and the host reads the right answer afterwards."""
    n = (1 << 21) + 3
    ha, hb, hc = (rng.normal(size=n).astype(np.float32) for _ in range(3))
    a, b, c = pn.managed_array(ha), pn.managed_array(hb), pn.managed_array(hc)
    t, u = pn.managed_empty(n, np.float32), pn.managed_empty(n, np.float32)
    pn.cuda_reset_stats()
    np.multiply(a, b, out=t)
    np.add(t, c, out=u)
    s = u.sum()
    m = u.max()
    i = u.argmin()
    np.add(u, np.float32(2.5), out=t)      # host scalar next to resident arrays
    info = pn.cuda_info()
    assert info["calls"] == 6 and info["h2d_bytes"] == 0 and info["staged_bytes"] == 0, info
    assert info["d2h_bytes"] <= 64          # only the reduction scalars came back
    g = np.greater(t, c)                    # output allocated by NumPy (pageable): inputs in place, result staged back
    info = pn.cuda_info()
    assert info["calls"] == 7 and info["h2d_bytes"] == 0 and info["d2h_bytes"] <= 64 + n, info
    want_u = ha * hb + hc
    assert_bits_equal(np.asarray(u), want_u, "chain on managed arrays")
    assert_bits_equal(np.asarray(t), want_u + np.float32(2.5), "scalar operand")
    assert np.array_equal(g, (want_u + np.float32(2.5)) > hc)
    assert m == want_u.max() and i == want_u.argmin()
    ref = float(np.sum(want_u.astype(np.float64)))
    assert abs(float(s) - ref) <= 2.0 ** -23 * abs(ref) + 2.0 ** -45 * float(np.abs(want_u).sum())


def test_large_managed_arrays_shard_over_all_gpus(pn, rng):
    """On a multi-GPU box a large managed ndarray is spread over the GPUs (shard l prefetched to GPU l) and every
    call runs one kernel per GPU in place: same bits as one GPU, no PCIe, `gpus_last` == the device count.
    On one GPU the same calls take the single-GPU resident path (checked for the same results)."""
    n = (1 << 25) + 12345   # 128 MiB float32: above the 64 MiB sharding threshold
    ha, hb = rng.normal(0, 10, n).astype(np.float32), rng.normal(0, 10, n).astype(np.float32)
    ha[n // 3] = np.float32(1e6)
    a, b = pn.managed_array(ha), pn.managed_array(hb)
    t, u = pn.managed_empty(n, np.float32), pn.managed_empty(n, np.float32)
    ndev = pn.cuda_info()["devices"]
    pn.cuda_reset_stats()
    for _ in range(2):   # second round: the shards are already where they belong
        np.multiply(a, b, out=t)
        np.add(t, np.float32(1.5), out=u)
        np.sin(u, out=t)
        u += a
        s, m, i, k = u.sum(), u.max(), int(u.argmax()), int(u.argmin())
    info = pn.cuda_info()
    assert info["h2d_bytes"] == 0 and info["staged_bytes"] == 0, info
    assert info["gpus_last"] == ndev and info["lanes_last"] == ndev, info
    want_u = ha * hb + np.float32(1.5)
    pn.disable()
    want_t = None
    try:
        from pnumpy_b200 import device as dev
        want_t = dev.unary("SIN", dev.to_device(want_u)).to_host()   # the GPU body's bits (Cephes), on one GPU
    finally:
        pn.enable()
    want_u2 = want_u + ha
    assert_bits_equal(np.asarray(t), want_t, "sin on sharded managed arrays")
    assert_bits_equal(np.asarray(u), want_u2, "in place on sharded managed arrays")
    assert m == want_u2.max() and i == int(want_u2.argmax()) and k == int(want_u2.argmin())
    ref = float(np.sum(want_u2.astype(np.float64)))
    assert abs(float(s) - ref) <= 2.0 ** -23 * abs(ref) + 2.0 ** -45 * float(np.abs(want_u2).sum())
    # the sum does not depend on how many GPUs hold the array: compare with one GPU's two passes
    assert np.float32(s).tobytes() == dev.reduce("ADD", dev.to_device(want_u2)).to_host().tobytes()
    # the fused chain follows the shards too
    pn.cuda_reset_stats()
    f = pn.muladd(a, b, u)
    fs = pn.muladd_sum(a, b, u)
    info = pn.cuda_info()
    assert info["calls"] == 2 and info["h2d_bytes"] == 0 and info["gpus_last"] == ndev, info
    assert_bits_equal(np.asarray(f), ha * hb + want_u2, "muladd on sharded managed arrays")
    assert np.float32(fs).tobytes() == dev.reduce("ADD", dev.to_device(ha * hb + want_u2)).to_host().tobytes()


def test_recycler_gives_pinned_arrays(pn, rng):
    assert pn.recycler_isenabled() is False
    pn.recycler_enable()
    try:
        assert pn.recycler_isenabled() is True
        n = 1 << 21
        a = np.ones(n, np.float32)       # born pinned
        b = np.full(n, 2.0, np.float32)
        pn.cuda_reset_stats()
        c = a + b                        # output allocated by NumPy through the recycler too
        info = pn.cuda_info()
        assert info["calls"] == 1 and info["staged_bytes"] == 0
        assert np.all(c == 3)
        del c
        d = a * b                        # same size: the freed block is recycled
        assert pn.recycler_info()["hits"] >= 1 and np.all(d == 2)
        small = np.ones(10)              # small arrays stay on malloc
        assert small.sum() == 10
    finally:
        pn.recycler_disable()
    assert pn.recycler_isenabled() is False
    assert np.all(np.ones(1 << 21, np.float32) + 1 == 2)


def test_ledger_counts(pn):
    pn.ledger_enable()
    try:
        assert pn.ledger_isenabled()
        before = pn.ledger_info()["Binary"]["gpu_calls"]
        a = np.ones(1 << 16)
        a + a
        assert pn.ledger_info()["Binary"]["gpu_calls"] == before + 1
    finally:
        pn.ledger_disable()
    assert not pn.ledger_isenabled()


@pytest.mark.parametrize("dt", ["int8", "uint8", "int16", "uint16", "int32", "uint32", "int64", "uint64"])
def test_int_true_divide_loops(pn, rng, dt):
    """The added (intN, intN) -> float64 true_divide loops: operands reach the GPU uncast; the result is
    bit-for-bit NumPy's cast-then-divide, on the GPU path, on the NumPy path (atop disabled -> NumPy's own
    casts + its dd->d loop) and for declined layouts."""
    assert pn.cuda_info()["extra_loops"] == 10
    n = (1 << 20) + 5
    a, b = fill_random(n, dt, rng), fill_random(n, dt, rng)
    b[::7] = 0
    a[::21] = 0  # 0/0 -> nan as well as x/0 -> inf
    with np.errstate(all="ignore"):
        want = a.astype(np.float64) / b.astype(np.float64)
        pn.cuda_reset_stats()
        got = np.true_divide(a, b)
        info = pn.cuda_info()
        assert info["calls"] == 1 and info["h2d_bytes"] == 2 * a.nbytes, "operands must travel uncast"
        assert_bits_equal(got, want, f"gpu {dt}")
        assert_bits_equal(a / b[3], a.astype(np.float64) / np.float64(b[3]), "scalar divisor")
        assert_bits_equal(a[::2] / b[::2], want[::2].copy(), "strided -> NumPy path")
        out = np.empty(n, np.float64)
        np.divide(a, b, out=out)
        assert_bits_equal(out, want, "out=")
        pn.disable()
        try:
            assert_bits_equal(np.true_divide(a, b), want, f"numpy path {dt}")
            assert_bits_equal(np.true_divide(a[:1000:3], b[:1000:3]), want[:1000:3].copy(), "numpy path strided")
        finally:
            pn.enable()
        # mixed widths still promote through NumPy and land on one of the added loops or dd->d
        other = b.astype(np.int64 if np.dtype(dt).kind == "i" else np.uint64)
        assert_bits_equal(np.true_divide(a, other), a.astype(np.float64) / other.astype(np.float64), "mixed")


def test_benchmark_plumbing(pn, capsys):
    """config 1 of BASELINE.json: pn.benchmark plumbing on 1 000 000 float32 elements."""
    old = pn.cuda_setthreshold(1 << 16)
    try:
        r = pn.benchmark_func(np.add, ctypes=[np.float32], sizes=[1_000_000], loop_size=5)
        assert r.shape == (1,) and np.isfinite(r[0]) and r[0] > 0
        assert pn.atop_isenabled() and pn.thread_isenabled()
        pn.benchmark(ctypes=[np.float32], sizes=[1 << 17], loop_size=3)
        out = capsys.readouterr().out
        assert out.startswith("131072 rows,float32,") and "a+b," in out and "sum," in out and "min," in out
        pn.benchmark(sizes=[1 << 16], loop_size=2)          # the default dtype list (bool, int8 ... float64) must run
        out = capsys.readouterr().out
        assert out.startswith("65536 rows,bool,int8,int16,int32,int64,float32,float64,") and out.count("\n") >= 12
    finally:
        pn.cuda_setthreshold(old)


# ---- fused chain: pn.muladd / pn.muladd_sum ---------------------------------------------
@pytest.mark.parametrize("dt", ["float32", "float64", "int32", "int64"])
def test_fused_chain_equals_the_hooked_chain(pn, rng, dt):
    """pn.muladd / pn.muladd_sum == np.add(np.multiply(a, b), c) [.sum()] bit for bit -- against NumPy's own
    loops element-wise, against the three hooked GPU calls for the sum -- with 3 operands in, 1 (or none) out
    over PCIe, for pageable and pinned arrays and any number of lanes."""
    n = (1 << 22) + 4099
    if np.dtype(dt).kind == "f":
        ha, hb, hc = (rng.normal(0, 100, n).astype(dt) for _ in range(3))
    else:
        ha, hb, hc = (fill_random(n, dt, rng) for _ in range(3))
    isz = np.dtype(dt).itemsize
    pn.disable()
    want = np.add(np.multiply(ha, hb), hc)
    pn.enable()
    chain_sum = np.add(np.multiply(ha, hb), hc).sum()    # three hooked calls on the GPU
    oldw = pn.thread_getworkers()
    try:
        for workers in (1, 3, 7):
            pn.thread_setworkers(workers)
            pn.cuda_reset_stats()
            got = pn.muladd(ha, hb, hc)
            info = pn.cuda_info()
            assert info["calls"] == 1 and info["h2d_bytes"] == 3 * n * isz and info["d2h_bytes"] == n * isz, info
            assert_bits_equal(got, want, f"pn.muladd {dt} workers={workers}")
            pn.cuda_reset_stats()
            s = pn.muladd_sum(ha, hb, hc)
            info = pn.cuda_info()
            if dt == "int32":   # ndarray.sum() widens int32 to int64: pn.muladd_sum runs the three hooked calls
                assert info["calls"] == 3, info
            else:
                assert info["calls"] == 1 and info["h2d_bytes"] == 3 * n * isz and info["d2h_bytes"] <= 64, info
            assert type(s) is type(chain_sum) and s.tobytes() == chain_sum.tobytes(), (dt, workers, s, chain_sum)
    finally:
        pn.thread_setworkers(oldw)
    # pinned operands: DMA'd in place, nothing staged; out= honoured and may alias an input
    pa, pb, pc = pn.pinned_array(ha), pn.pinned_array(hb), pn.pinned_array(hc)
    pn.cuda_reset_stats()
    r = pn.muladd(pa, pb, pc, out=pc)
    assert r is pc and pn.cuda_info()["staged_bytes"] == 0 and pn.cuda_info()["calls"] == 1
    assert_bits_equal(np.asarray(pc), want, f"pn.muladd {dt} pinned, in place")


def test_fused_chain_on_resident_arrays_and_fallbacks(pn, rng):
    n = (1 << 21) + 3
    ha, hb, hc = (rng.normal(size=n).astype(np.float32) for _ in range(3))
    want = ha * hb + hc
    a, b, c = pn.managed_array(ha), pn.managed_array(hb), pn.managed_array(hc)
    pn.cuda_reset_stats()
    u = pn.muladd(a, b, c)                   # result allocated in managed memory too
    s = pn.muladd_sum(a, b, c)
    info = pn.cuda_info()
    assert info["calls"] == 2 and info["h2d_bytes"] == 0 and info["staged_bytes"] == 0 and info["d2h_bytes"] <= 64, info
    assert_bits_equal(np.asarray(u), want, "muladd on managed arrays")
    assert s.tobytes() == np.add(np.multiply(a, b), c).sum().tobytes()
    # operands that do not qualify run the ordinary ufunc chain: same values, NumPy's broadcasting and casting
    pn.cuda_reset_stats()
    r = pn.muladd(ha[::2], hb[::2], hc[::2])                      # non-contiguous
    assert_bits_equal(r, want[::2], "strided fallback")
    r = pn.muladd(ha, np.float32(3.0), hc)                        # scalar operand
    assert_bits_equal(r, ha * np.float32(3.0) + hc, "scalar fallback")
    r = pn.muladd(ha, hb.astype(np.float64), hc)                  # mixed dtypes
    assert r.dtype == np.float64 and np.array_equal(r, ha * hb.astype(np.float64) + hc)
    assert pn.muladd_sum(ha[::2], hb[::2], hc[::2]).tobytes() == np.add(np.multiply(ha[::2], hb[::2]), hc[::2]).sum().tobytes()
    assert pn.muladd_sum([1.0, 2.0], [3.0, 4.0], [5.0, 6.0]) == 22.0
    pn.disable()
    try:
        assert_bits_equal(pn.muladd(ha, hb, hc), want, "atop disabled -> NumPy chain")
    finally:
        pn.enable()


@pytest.mark.parametrize("dt", FP_DTYPES)
def test_more_transcendental_ufuncs_run_on_the_gpu(pn, rng, dt):
    """tan ... cbrt (hooked by the reference without a body; it threads NumPy's loop): here the hooked call
    reaches a GPU body, within 4 ULP of NumPy's own loop on the same input."""
    n = 1 << 20
    u = rng.uniform(-0.99, 0.99, n).astype(dt)
    cases = {"tan": u * 10, "arcsin": u, "arccos": u, "arctan": u * 50, "sinh": u * 10, "cosh": u * 10, "tanh": u * 5,
             "arcsinh": u * 1e3, "arccosh": np.abs(u) * 100 + 1, "arctanh": u, "log2": np.abs(u) * 1e3 + 1e-3,
             "log10": np.abs(u) * 1e3 + 1e-3, "exp2": u * 60, "expm1": u * 10, "log1p": u * 100 + 99, "cbrt": u * 1e5}
    for name, v in cases.items():
        v = v.astype(dt)
        got, want, calls = _both(pn, lambda: getattr(np, name)(v))
        assert calls >= 1, name
        assert got.dtype == want.dtype
        assert ulp_error(got, want.astype(np.float64)).max() <= 4.0, (name, dt)


def test_concurrent_python_threads(pn, rng):
    """NumPy releases the GIL around these loops, so two Python threads can be inside the hooked loop at
    once.  The reference degrades the second caller to its unthreaded path (src/atop/threads.h:1064-1073);
    here nte's try_lock hands it NumPy's own loop.  Either way every result is right and nothing hangs."""
    import threading
    n = (1 << 22) + 3
    data = [(rng.normal(size=n).astype(np.float32), rng.normal(size=n).astype(np.float32)) for _ in range(4)]
    want = [(a + b, np.sin(a), a.sum(dtype=np.float32), int(np.argmax(b))) for a, b in data]
    got = [None] * 4
    errs = []

    def work(i):
        try:
            a, b = data[i]
            for _ in range(5):
                got[i] = (a + b, np.sin(a), a.sum(), int(b.argmax()))
        except Exception as e:  # pragma: no cover
            errs.append(e)
    c0, d0 = pn.cuda_info()["calls"], pn.cuda_info()["declined"]
    ths = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    for t in ths:
        t.start()
    for t in ths:
        t.join(timeout=120)
    assert not any(t.is_alive() for t in ths), "a hooked call hung"
    assert not errs, errs
    info = pn.cuda_info()
    assert info["calls"] > c0  # some calls ran on the GPU ...
    for i in range(4):
        assert_bits_equal(got[i][0], want[i][0], "add under contention")
        # sin: the GPU body (Cephes) and NumPy's differ in the last place; whichever ran is within 2 ULP here
        assert np.all(ulp_error(got[i][1], np.sin(data[i][0].astype(np.float64)).astype(np.float32)) <= 2)
        assert abs(float(got[i][2]) - float(want[i][2])) <= 2.0 ** -20 * np.abs(data[i][0]).sum()
        assert got[i][3] == want[i][3]
    assert info["declined"] >= d0


def test_degenerate_sizes(pn):
    """empty, one-element and just-below/above-a-vector arrays through the hooks (threshold 0: all on the GPU)"""
    for n in (0, 1, 2, 7, 8, 9, 31, 32, 33, 255, 256, 257):
        for dt in ("int8", "int32", "float32", "float64"):
            a = (np.arange(n) % 5 + 1).astype(dt)
            b = (np.arange(n) % 3 + 1).astype(dt)
            got, want, _ = _both(pn, lambda: (a + b, a * b, a > b, np.abs(a), a.sum() if n else 0, a.max() if n else 0,
                                               int(a.argmax()) if n else 0))
            for g, w in zip(got, want):
                assert np.asarray(g).tobytes() == np.asarray(w).tobytes(), (n, dt)
    with pytest.raises(ValueError):
        np.empty(0, np.float32).argmax()
    with pytest.raises(ValueError):
        np.empty(0, np.float32).max()


@pytest.mark.parametrize("dt", ["int8", "int16", "int32", "float32", "float64"])
def test_host_shapes_around_shard_and_slab_boundaries(pn, rng, dt):
    """lengths around the 65 536-element shard unit (1..5 lanes' worth), host pointers at odd element offsets,
    1, 2 and 4 lanes: element-wise results bit-exact, reductions exact or in tolerance, against NumPy's loops."""
    unit = 65536
    base_a = rng.normal(0, 100, 5 * unit + 600).astype(dt)
    base_b = (rng.normal(0, 100, 5 * unit + 600) + 1).astype(dt)
    oldw = pn.thread_getworkers()
    try:
        for workers, threading in ((1, False), (1, True), (3, True)):
            pn.thread_enable() if threading else pn.thread_disable()
            pn.thread_setworkers(workers)
            for n in (unit - 1, unit, unit + 1, 2 * unit + 33, 4 * unit, 4 * unit + 511, 5 * unit + 257):
                for off in (0, 1, 5):
                    a, b = base_a[off:off + n], base_b[off + 2:off + 2 + n]
                    got, want, calls = _both(pn, lambda: (a + b, a * b, a > b, np.abs(a), a.max(), a.min(), int(a.argmax()), int(a.argmin())))
                    assert calls >= 8, (n, off, calls)
                    for g, w in zip(got, want):
                        assert np.asarray(g).tobytes() == np.asarray(w).tobytes(), (dt, n, off, workers)
                    s_got, s_want, _ = _both(pn, lambda: a.sum())
                    if np.dtype(dt).kind == "f":
                        tol = (2.0 ** -22 if dt == "float32" else 2.0 ** -48) * float(np.abs(a.astype(np.float64)).sum())
                        assert abs(float(s_got) - float(s_want)) <= tol, (dt, n, off)
                    else:
                        assert s_got == s_want
    finally:
        pn.thread_enable()
        pn.thread_setworkers(oldw)


def test_other_ways_numpy_calls_a_hooked_loop(pn, rng):
    """accumulate (the output of element i is the input of element i+1: an overlapping call the dispatcher must
    hand back), where= masks, ufunc.at, outer, reduceat, broadcasting and keepdims/axis reductions: whatever
    argument shapes NumPy builds, the results are NumPy's."""
    n = (1 << 20) + 11
    a = rng.normal(0, 3, n).astype(np.float64)
    b = rng.normal(0, 3, n).astype(np.float64)
    ai = rng.integers(-50, 50, n).astype(np.int32)
    m = rng.random(n) > 0.3
    idx = rng.integers(0, 1000, 50000)
    cases = {
        "cumsum": lambda: np.cumsum(a),
        "cumsum int": lambda: np.add.accumulate(ai),
        "maximum.accumulate": lambda: np.maximum.accumulate(a),
        "where": lambda: np.add(a, b, out=np.zeros(n), where=m),
        "where sqrt": lambda: np.sqrt(np.abs(a), out=np.full(n, -1.0), where=m),
        "at": lambda: _at(np.zeros(1000), idx),
        "outer": lambda: np.multiply.outer(a[:1500], b[:1200]),
        "reduceat": lambda: np.add.reduceat(ai, np.arange(0, n, 70001)),
        "broadcast row": lambda: a.reshape(-1, 1)[:4096] * b[:2048],
        "broadcast scalar arr": lambda: a.reshape(1, -1) + b[:1].reshape(1, 1),
        "axis0": lambda: a[: 1 << 20].reshape(16, -1).sum(axis=0),
        "axis1 keepdims": lambda: a[: 1 << 20].reshape(16, -1).max(axis=1, keepdims=True),
        "sum initial": lambda: np.add.reduce(ai, initial=7),
        "min initial": lambda: np.minimum.reduce(a, initial=-1e9),
        "sum where": lambda: np.add.reduce(a, where=m),
        "sum dtype": lambda: a.astype(np.float32).sum(dtype=np.float64),
        "divmod path": lambda: np.floor_divide(a, b),
        "comparison chain": lambda: (a > b) & (a < 2 * b) | (ai == 3),
    }
    for name, fn in cases.items():
        got, want, _ = _both(pn, fn)
        g, w = np.asarray(got), np.asarray(want)
        assert g.dtype == w.dtype and g.shape == w.shape, name
        if g.dtype.kind == "f" and ("sum" in name or name in ("axis0",)):
            assert np.allclose(g, w, rtol=1e-12, atol=1e-9), name        # float sums: order of summation differs
        else:
            assert_bits_equal(g, w, name)


def _at(z, idx):
    np.add.at(z, idx, 1.0)
    return z


# ---- round 2: residency through the unmodified NumPy API, error paths, ledger records -------------------------
def test_managed_recycler_makes_numpy_arrays_device_capable(pn, rng):
    """recycler_enable(kind="managed"): the arrays NumPy itself creates -- np.empty, ufunc outputs, the temporaries
    of a*b+c -- live in CUDA managed memory, so a plain NumPy expression runs on the GPU with no PCIe copies, at
    every size, and only the reduction scalar comes back (reference roadmap: src/pnumpy/__init__.py:18-21,
    doc_src/source/roadmap.rst:27-31)."""
    from pnumpy_b200 import _pnumpy
    old_thr = pn.cuda_setthreshold(1 << 19)      # the DEFAULT host threshold: residency must not depend on it
    pn.recycler_enable(kind="managed")
    try:
        assert pn.recycler_info()["kind"] == "managed"
        n = 1_000_000                             # pn.benchmark's size: far below the host-operand threshold
        a = (np.arange(n, dtype=np.int64) % 253 + 1).astype(np.float32)   # astype output: born managed
        b = a.copy()
        assert _pnumpy.pointer_kind(a.ctypes.data) == 3 and _pnumpy.pointer_kind(b.ctypes.data) == 3
        tiny = np.ones(100, np.float32)
        assert _pnumpy.pointer_kind(tiny.ctypes.data) == 0      # small arrays stay on malloc
        pn.cuda_reset_stats()
        c = a * b + a                             # two ufunc calls, one temporary: all managed, all on the GPU
        s = c.sum()
        eq = a == b                               # 1 000 000-byte bool output: above the managed threshold too
        info = pn.cuda_info()
        assert info["calls"] == 4 and info["declined"] == 0, info
        assert info["h2d_bytes"] == 0 and info["staged_bytes"] == 0 and info["d2h_bytes"] <= 64, info
        assert _pnumpy.pointer_kind(c.ctypes.data) == 3 and _pnumpy.pointer_kind(eq.ctypes.data) == 3
        ha = (np.arange(n, dtype=np.int64) % 253 + 1).astype(np.float64)
        assert s.dtype == np.float32 and s == np.float32(np.sum(ha * ha + ha))   # float64 accumulation, rounded once
        assert eq.all()
        # steady state: nothing is prefetched again, recycled blocks are reused
        np.add(a, b, out=c)
        d = a + b                                 # a fresh block is placed once; recycled, it stays where it is
        del d
        pre = pn.cuda_info()["prefetches"]
        for _ in range(5):
            np.add(a, b, out=c)
            d = a + b
            del d
        assert pn.cuda_info()["prefetches"] == pre
        assert pn.recycler_info()["hits"] >= 4
        z = np.zeros(n, np.float32)               # zeroed on the GPU
        assert _pnumpy.pointer_kind(z.ctypes.data) == 3 and not z.any()
    finally:
        pn.recycler_disable()
        pn.cuda_setthreshold(old_thr)
        pn.cuda_setthreshold(0)
    assert _pnumpy.pointer_kind(np.ones(n, np.float32).ctypes.data) == 0


def test_small_ops_on_managed_memory_stay_on_the_gpu(pn, rng):
    """A small op on (a slice of) a managed array must not fall to NumPy's CPU loop (which would migrate the pages
    back to the host): resident operands run on the GPU at EVERY size, whatever the host-operand threshold."""
    old_thr = pn.cuda_setthreshold(1 << 19)
    pn.ledger_enable()
    pn.ledger_clear()
    try:
        h = rng.normal(size=1 << 20).astype(np.float32)
        m = pn.managed_array(h)
        out = pn.managed_empty(1 << 20, np.float32)
        np.add(m, m, out=out)                     # places the pages
        cnts = (2, 7, 1000, 4097, 65537)
        pn.disable()                              # expectations from NumPy's own loops, on the host copy
        want = [(float(np.sum(h[100:100 + cnt], dtype=np.float64)), int(h[:cnt].argmax())) for cnt in cnts]
        pn.enable()
        pn.cuda_reset_stats()
        pn.ledger_clear()
        for cnt, (wsum, warg) in zip(cnts, want):
            np.multiply(m[5:5 + cnt], m[9:9 + cnt], out=out[3:3 + cnt])
            assert float(m[100:100 + cnt].sum()) == pytest.approx(wsum, rel=1e-6)
            assert int(m[:cnt].argmax()) == warg
        info = pn.cuda_info()
        assert info["calls"] == 15 and info["declined"] == 0 and info["h2d_bytes"] == 0, info
        li = pn.ledger_info()
        assert li["Binary"]["numpy_calls"] == 0 and li["ArgMinMax"]["numpy_calls"] == 0, li   # no CPU loop touched the pages
        assert_bits_equal(np.asarray(out[3:3 + 65537]), h[5:5 + 65537] * h[9:9 + 65537], "small slices")
        # strided managed views run on the GPU too (strided kernels), reductions included
        pn.cuda_reset_stats()
        k = h[1::3].size
        np.add(m[::3][:k], m[1::3], out=out[:k])
        r = float(m[::2].max())
        assert pn.cuda_info()["calls"] == 2 and pn.cuda_info()["declined"] == 0
        pn.disable()
        assert r == float(h[::2].max())
        assert_bits_equal(np.asarray(out[:k]), h[::3][:k] + h[1::3], "strided managed views")
    finally:
        pn.enable()
        pn.ledger_disable()
        pn.cuda_setthreshold(old_thr)
        pn.cuda_setthreshold(0)


def test_accumulate_and_partial_overlap_on_managed_arrays(pn, rng):
    """NumPy's accumulate calls a binary loop as (out, in + s, out + s): partially overlapping operands need the
    sequential semantics of NumPy's own loop.  The dispatcher must DECLINE that for resident operands exactly as
    it does for host ones -- not raise."""
    n = 300_000
    h = rng.integers(-50, 50, n).astype(np.int64)
    m, m2 = pn.managed_array(h), pn.managed_empty(n, np.int64)
    np.cumsum(m, out=m2)
    assert np.array_equal(np.asarray(m2), np.cumsum(h))
    np.add.accumulate(m, out=m)
    assert np.array_equal(np.asarray(m), np.cumsum(h))
    f = pn.managed_array(rng.normal(size=n))
    want = np.maximum.accumulate(np.asarray(f).copy())
    np.maximum.accumulate(f, out=f)
    assert np.array_equal(np.asarray(f), want)
    g = pn.managed_array(h.astype(np.float64))
    hg = h.astype(np.float64)
    np.add(g[1:], g[:-1], out=g[1:])             # exact aliasing of out with in1: allowed, parallel-safe
    hg[1:] = hg[1:] + hg[:-1].copy()
    assert np.array_equal(np.asarray(g), hg)


def test_mixed_resident_and_host_operands(pn, rng):
    """A call that mixes a managed array with ordinary host arrays: the resident operand is used in place, the host
    ones are staged; nothing falls back to the CPU loop and the bytes moved are those of the host operands only."""
    n = (1 << 21) + 77
    ha, hb = rng.normal(size=n), rng.normal(size=n)
    ma = pn.managed_array(ha)
    out = np.empty(n)
    pn.cuda_reset_stats()
    np.subtract(ma, hb, out=out)                  # managed - pageable -> pageable
    info = pn.cuda_info()
    assert info["calls"] == 1 and info["declined"] == 0 and info["h2d_bytes"] == 8 * n and info["d2h_bytes"] == 8 * n, info
    assert_bits_equal(out, ha - hb, "managed - host")
    mo = pn.managed_empty(n, np.float64)
    pn.cuda_reset_stats()
    np.multiply(hb, ha, out=mo)                   # host * host -> managed: no D2H
    info = pn.cuda_info()
    assert info["calls"] == 1 and info["h2d_bytes"] == 16 * n and info["d2h_bytes"] == 0, info
    assert_bits_equal(np.asarray(mo), hb * ha, "host * host -> managed")
    g = np.greater(ma, 0.25)                      # managed vs host scalar -> pageable bool
    assert np.array_equal(g, ha > 0.25)


def test_same_operand_is_uploaded_once(pn, rng):
    """np.multiply(a, a): the reference loads the operand once when in1 == in2
    (src/atop/ops_binary.cpp:631-658); over PCIe that halves the H2D bytes."""
    n = 1 << 22
    a = pn.pinned_array(rng.normal(size=n).astype(np.float32))
    out = pn.pinned_empty(n, np.float32)
    pn.cuda_reset_stats()
    np.multiply(a, a, out=out)
    info = pn.cuda_info()
    assert info["calls"] == 1 and info["h2d_bytes"] == 4 * n and info["d2h_bytes"] == 4 * n, info
    assert_bits_equal(np.asarray(out), np.asarray(a) * np.asarray(a), "a*a")
    pn.cuda_reset_stats()
    np.add(a, a[::-1], out=out)                   # same buffer, different view: two uploads
    assert pn.cuda_info()["h2d_bytes"] == 8 * n
    assert_bits_equal(np.asarray(out), np.asarray(a) + np.asarray(a)[::-1], "a + a[::-1]")
    r = pn.muladd(a, a, a)                        # the fused chain uploads a repeated operand once as well
    assert_bits_equal(np.asarray(r), np.asarray(a) * np.asarray(a) + np.asarray(a), "muladd(a, a, a)")


def test_cuda_failure_raises_runtime_error_and_the_process_stays_usable(pn, rng):
    """SURVEY.md 8(b) error convention, verified on this NumPy: a CUDA failure inside a hooked loop (injected here:
    the dispatcher reports cudaErrorLaunchFailure without touching the device) surfaces as RuntimeError from the
    NumPy call itself -- never a silently wrong array, never NumPy's CPU loop -- and the next call works."""
    n = 1 << 20
    a, b = rng.normal(size=n), rng.normal(size=n)
    m = pn.managed_array(a)
    for what, call in (("binary, host operands", lambda: np.add(a, b)),
                       ("unary", lambda: np.sqrt(np.abs(a))),
                       ("compare", lambda: np.less(a, b)),
                       ("reduction", lambda: a.sum()),
                       ("arg-reduction", lambda: a.argmax()),
                       ("binary, resident operands", lambda: np.add(m, m)),
                       ("resident reduction", lambda: m.max()),
                       ("int true_divide", lambda: np.true_divide(np.arange(1, n + 1), np.arange(2, n + 2)))):
        assert pn.cuda_inject_fault(1) == 0
        with pytest.raises(RuntimeError, match="failed on the GPU"):
            call()
        assert pn.cuda_inject_fault(0) == 0     # consumed
        got, want, calls = _both(pn, call)      # the very next call runs on the GPU and is right
        assert calls >= 1, what
        if isinstance(got, np.ndarray):
            assert np.array_equal(got, want), what
        else:
            assert got == pytest.approx(want, rel=1e-12), what
    # the fused entry points report through their own return path
    pn.cuda_inject_fault(1)
    with pytest.raises(RuntimeError, match="failed on the GPU"):
        pn.muladd_sum(a, b, a)
    assert float(pn.muladd_sum(a, b, a)) == pytest.approx(float(np.sum(a * b + a)), rel=1e-9)


def test_forked_child_runs_numpy_loops(pn, rng):
    """CUDA cannot be used in a fork()ed child.  The child must not hang on worker threads that do not exist there
    nor raise from dead CUDA calls: every hooked call is declined and NumPy's own loop runs (fork-start
    multiprocessing workers that use NumPy keep working, on the CPU)."""
    n = 1 << 21
    a, b = rng.normal(size=n), rng.normal(size=n)
    want = a + b
    (a + b)                                     # make sure the parent has lanes and worker threads
    r, w = os.pipe()
    pid = os.fork()
    if pid == 0:
        code = 1
        try:
            import signal
            signal.alarm(60)                    # a hang is a failure, not a stuck test run
            ok = np.array_equal(a + b, want) and float(a.sum()) == pytest.approx(float(want.sum() - b.sum()), rel=1e-9)
            ok = ok and int(a.argmax()) == int(np.argmax(a)) and pn.cuda_info()["forked_child"] == 1
            big = np.empty(1 << 22)             # allocation paths must not touch CUDA either
            big[:] = 1.0
            ok = ok and float(big.sum()) == float(1 << 22)
            code = 0 if ok else 2
        except BaseException:
            code = 3
        os.write(w, bytes([code]))
        os._exit(code)
    os.close(w)
    _, status = os.waitpid(pid, 0)
    got = os.read(r, 1)
    os.close(r)
    assert os.WIFEXITED(status) and os.WEXITSTATUS(status) == 0 and got == b"\x00", (status, got)
    got, want2, calls = _both(pn, lambda: a + b)  # the parent is unaffected
    assert calls == 1 and np.array_equal(got, want2)


def test_ledger_records_each_call(pn, rng):
    """reference: the ledger ring of per-call records (src/pnumpy/ledger.cpp:144-272)."""
    a = rng.normal(size=1 << 18).astype(np.float32)
    pn.ledger_enable()
    pn.ledger_clear()
    try:
        np.add(a, a)
        np.sin(a)
        a.sum()
        a.argmax()
        pn.disable()
        np.add(a, a)
        pn.enable()
    finally:
        pn.ledger_disable()
    recs = pn.ledger_records()
    assert [r["category"] for r in recs] == ["Binary", "TrigLog", "Binary", "ArgMinMax", "Binary"], recs
    assert [r["gpu"] for r in recs] == [True, True, True, True, False]
    assert recs[0]["elements"] == a.size and recs[0]["bytes"] == 12 * a.size and recs[0]["reduce"] is False
    assert recs[2]["reduce"] is True and recs[2]["bytes"] == 4 * (a.size - 1) or recs[2]["bytes"] == 4 * a.size
    assert all(r["end_tsc"] > r["start_tsc"] > 0 and r["ticks"] == r["end_tsc"] - r["start_tsc"] for r in recs)
    assert len(pn.ledger_records(2)) == 2 and pn.ledger_records(2)[-1] == recs[-1]


def test_calling_threads_device_is_left_alone(pn, rng):
    """The dispatcher drives every GPU from the calling thread; afterwards the thread's current CUDA device must be
    what it was (torch / cupy code in the same thread relies on it)."""
    from pnumpy_b200 import cabi
    lib = cabi.load()
    import ctypes
    if pn.cuda_info()["devices"] < 2:
        pytest.skip("needs two GPUs to be meaningful")
    cabi.check(lib.pncu_set_device(1), "set_device")
    try:
        a = pn.managed_array(rng.normal(size=1 << 25).astype(np.float32))
        (a + a).sum()
        np.add(rng.normal(size=1 << 21), 1.0)
        d = ctypes.c_int(-1)
        cabi.check(lib.pncu_get_device(ctypes.byref(d)), "get_device")
        assert d.value == 1
    finally:
        cabi.check(lib.pncu_set_device(0), "set_device")


def test_astype_through_the_dispatcher(pn, rng):
    """pn.astype: dtype conversion (the reference's roadmap item) through nte -- host arrays staged, managed arrays in
    place -- equal to NumPy's astype; what the GPU path does not take falls back to NumPy's own cast."""
    n = (1 << 20) + 5
    h = rng.normal(0, 1000, n)
    for src_dt, dst_dt in (("float64", "float32"), ("float64", "int32"), ("int64", "float64"), ("float32", "int8"), ("int16", "bool"),
                           ("uint8", "float32"), ("bool", "int64")):
        a = h.astype(src_dt)
        pn.cuda_reset_stats()
        got = pn.astype(a, dst_dt)
        assert pn.cuda_info()["calls"] == 1, (src_dt, dst_dt)
        with np.errstate(all="ignore"):
            assert_bits_equal(got, a.astype(dst_dt), f"{src_dt} -> {dst_dt}")
    m = pn.managed_array(h.astype(np.float32))
    pn.cuda_reset_stats()
    r = pn.astype(m, np.float64)
    info = pn.cuda_info()
    from pnumpy_b200 import _pnumpy
    assert info["calls"] == 1 and info["h2d_bytes"] == 0 and info["d2h_bytes"] == 0 and _pnumpy.pointer_kind(r.ctypes.data) == 3
    assert_bits_equal(np.asarray(r), h.astype(np.float32).astype(np.float64), "managed float32 -> float64")
    out = np.empty(n, np.int16)
    assert pn.astype(h, np.int16, out=out) is out
    with np.errstate(all="ignore"):
        assert_bits_equal(out, h.astype(np.int16), "out=")
    c = np.arange(12.0).reshape(3, 4)[:, ::2]           # non-contiguous: NumPy's cast
    assert np.array_equal(pn.astype(c, np.int32), c.astype(np.int32))
    z = np.arange(5.0).astype(np.complex128)             # no kernel: NumPy's cast
    assert np.array_equal(pn.astype(z, np.complex64), z.astype(np.complex64))
