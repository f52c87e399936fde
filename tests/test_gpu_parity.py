"""GPU parity tests proper: the sm_100a kernels, called through the C ABI (ctypes on
libpnumpy_cuda.so, device pointers), against the CPU oracle on the same seeded inputs, against
the golden fixtures, and -- at the BASELINE.json sizes -- through size-independent properties.

Bar (BASELINE.json north_star): bit-exact for integer / bool / comparison / min / max / arg* and
IEEE float add, sub, mul, div, sqrt; float32 sin/cos/log/exp bit-exact with the oracle's Cephes
restatement (hence 0 ULP from the reference's vector body on its valid domain) and <= 2 ULP of the
correctly rounded value elsewhere; float64 sin/cos/log/exp <= 2 ULP; float sums within
    |gpu - S64| <= 2^-23 |S64| + 2^-50 sum|x|      (float32)
    |gpu - S|   <= (log2(n) + 2) 2^-53 sum|x|      (float64)
"""
import math
import zlib

import numpy as np
import pytest

from inputs import fill_random, nan_case_input, reduction_input
from util import DTYPES, FP_DTYPES, INT_DTYPES, assert_bits_equal, ulp_error

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    from pnumpy_b200 import cabi, device
    cabi.init()  # raises without a B200: the tests must not pass on a fallback
    return device


@pytest.fixture(scope="module")
def O():
    from oracle import oracle
    return oracle


def _ops_for(dt):
    ops = ["ADD", "MUL", "MIN", "MAX"]
    if dt != "bool":
        ops.append("SUB")
    if dt in FP_DTYPES:
        ops += ["DIV", "FLOORDIV"]
    else:
        ops += ["BITWISE_AND", "BITWISE_OR", "BITWISE_XOR"]
    if dt in INT_DTYPES:
        ops += ["BITWISE_LSHIFT", "BITWISE_RSHIFT"]
    if dt == "bool":
        ops += ["LOGICAL_AND", "LOGICAL_OR"]
    return ops


N_SMALL = (1 << 16) + 13  # not a multiple of any vector width: exercises head/tail code


@pytest.mark.parametrize("dt", DTYPES)
def test_binary_vs_oracle(dev, O, rng, dt):
    launches0 = _launches()
    a, b = fill_random(N_SMALL, dt, rng), fill_random(N_SMALL, dt, rng)
    if dt in INT_DTYPES:
        b_shift = (b.astype(np.int64) % 70 - 3).astype(dt)  # shift counts around the width
    da, db = dev.to_device(a), dev.to_device(b)
    for op in _ops_for(dt):
        bb, dbb = (b_shift, dev.to_device(b_shift)) if "SHIFT" in op else (b, db)
        with np.errstate(all="ignore"):
            want = O.binary(op, a, bb)
        assert_bits_equal(dev.binary(op, da, dbb).to_host(), want, f"{op} {dt} vv")
        # scalar second operand and scalar first operand (stride 0)
        s = bb[7:8].copy()
        ds = dev.to_device(s)
        with np.errstate(all="ignore"):
            assert_bits_equal(dev.binary(op, da, ds, b_scalar=True).to_host(),
                              O.binary(op, a, np.broadcast_to(s, (1,))[0]), f"{op} {dt} vs")
            s1 = a[3:4].copy()
            assert_bits_equal(dev.binary(op, dev.to_device(s1), dbb, a_scalar=True).to_host(),
                              O.binary(op, np.broadcast_to(s1, (1,))[0], bb), f"{op} {dt} sv")
    assert _launches() > launches0  # the CUDA kernels really ran


def _launches():
    from pnumpy_b200 import cabi
    return cabi.load().pncu_launch_count()


@pytest.mark.parametrize("dt", ["int8", "int32", "float32", "float64"])
def test_binary_layouts(dev, O, rng, dt):
    """in-place, misaligned heads, ragged tiny sizes, byte strides (incl. negative)."""
    from pnumpy_b200 import cabi
    for n in (0, 1, 2, 3, 5, 15, 17, 31, 33, 255, 1023, 1025, 4099):
        a, b = fill_random(n, dt, rng), fill_random(n, dt, rng)
        with np.errstate(all="ignore"):
            want = O.binary("ADD", a, b) if n else np.empty(0, dt)
        got = dev.binary("ADD", dev.to_device(a), dev.to_device(b)).to_host() if n else want
        assert_bits_equal(got, want, f"n={n}")
    n = 50021
    a, b = fill_random(n + 8, dt, rng), fill_random(n + 8, dt, rng)
    da, db, dout = dev.to_device(a), dev.to_device(b), dev.empty(n + 8, dt)
    with np.errstate(all="ignore"):
        for off in (1, 2, 3, 5):  # same misalignment on all three: peeled head
            got = dev.binary("MUL", da.view(off, n), db.view(off, n), out=dout.view(off, n)).to_host()
            assert_bits_equal(got, O.binary("MUL", a[off:off + n], b[off:off + n]), f"off={off}")
        # different misalignment per operand: element-wise fallback kernel
        got = dev.binary("SUB", da.view(1, n), db.view(2, n), out=dout.view(3, n)).to_host()
        assert_bits_equal(got, O.binary("SUB", a[1:1 + n], b[2:2 + n]), "mixed misalignment")
        # in place: out aliases in1 exactly
        want = O.binary("ADD", a, b)
        dev.binary("ADD", da, db, out=da)
        assert_bits_equal(da.to_host(), want, "in-place")
    # arbitrary byte strides through the raw launcher, incl. a negative one
    fn, _ = cabi.get_binary(cabi.BINARY_OPS["ADD"], cabi.atop_of(dt))
    a, b = fill_random(3 * 4000, dt, rng), fill_random(2 * 4000, dt, rng)
    da, db, dout = dev.to_device(a), dev.to_device(b), dev.empty(4000, dt)
    isz = np.dtype(dt).itemsize
    cabi.check(fn(da.ptr, db.ptr + (2 * 4000 - 1) * isz, dout.ptr, 4000, 3 * isz, -2 * isz, isz, None), "strided")
    with np.errstate(all="ignore"):
        assert_bits_equal(dout.to_host(), O.binary("ADD", a[::3], b[::-2][:4000].copy()), "strided")
    # partial overlap is refused, never computed wrongly
    rc = fn(da.ptr, db.ptr, da.ptr + isz, 1000, isz, isz, isz, None)
    assert rc == -3 and b"overlaps" in cabi.load().pncu_last_error()


@pytest.mark.parametrize("dt", DTYPES)
def test_compare_vs_oracle(dev, O, rng, dt):
    a, b = fill_random(N_SMALL, dt, rng), fill_random(N_SMALL, dt, rng)
    b[::3] = a[::3]  # plenty of equal pairs
    da, db = dev.to_device(a), dev.to_device(b)
    for op in ("EQ", "NE", "LT", "GT", "LTE", "GTE"):
        assert_bits_equal(dev.compare(op, da, db).to_host(), O.compare(op, a, b), f"{op} {dt}")
        s = b[5:6].copy()
        assert_bits_equal(dev.compare(op, da, dev.to_device(s), b_scalar=True).to_host(),
                          O.compare(op, a, np.broadcast_to(s, (1,))[0]), f"{op} {dt} scalar")


@pytest.mark.parametrize("dt", DTYPES)
def test_unary_vs_oracle(dev, O, rng, dt):
    a = fill_random(N_SMALL, dt, rng)
    da = dev.to_device(a)
    ops = ["ABS", "LOGICAL_NOT", "ISNAN", "ISINF", "ISFINITE"]
    if dt in FP_DTYPES:
        ops += ["NEGATIVE", "SIGN", "FLOOR", "CEIL", "TRUNC", "ROUND", "SQRT", "SIGNBIT"]
    else:
        ops += ["INVERT"]
        if dt != "bool":
            ops += ["NEGATIVE"]
        if dt in ("int8", "int16", "int32", "int64"):
            ops += ["SIGN"]
    for op in ops:
        with np.errstate(all="ignore"):
            want = O.unary(op, a)
        got = dev.unary(op, da).to_host()
        assert_bits_equal(got, want.view(got.dtype), f"{op} {dt}")


@pytest.mark.parametrize("dt", INT_DTYPES)
def test_int_true_divide(dev, O, rng, dt):
    a, b = fill_random(N_SMALL, dt, rng), fill_random(N_SMALL, dt, rng)
    b[::5] = 0
    with np.errstate(all="ignore"):
        want = O.int_true_divide(a, b)
        assert_bits_equal(want, np.true_divide(a.astype(np.float64), b.astype(np.float64)), "oracle==numpy")
    assert_bits_equal(dev.int_true_divide(dev.to_device(a), dev.to_device(b)).to_host(), want, dt)


# ---- transcendental ---------------------------------------------------------------------------
TRIG_RANGES = {"SIN": (-8192, 8192), "COS": (-8192, 8192), "LOG": (0, 1e6), "EXP": (-80, 80)}


@pytest.mark.parametrize("op", ["SIN", "COS", "LOG", "EXP"])
def test_trig_f32_bit_exact_vs_oracle(dev, O, rng, golden, op):
    lo, hi = TRIG_RANGES[op]
    x = rng.uniform(lo, hi, 1 << 18).astype(np.float32)
    with np.errstate(all="ignore"):
        assert_bits_equal(dev.unary(op, dev.to_device(x)).to_host(), O.unary(op, x), f"{op} range")
        # every bit pattern class: NaN, inf, denormals, huge
        y = fill_random(1 << 18, "float32", rng)
        got, want = dev.unary(op, dev.to_device(y)).to_host(), O.unary(op, y)
        if op in ("SIN", "COS"):
            m = np.abs(y) <= 8192  # Cephes restatement: bit-exact; beyond: CUDA vs glibc libm
            assert_bits_equal(got[m], want[m], f"{op} bits in-range")
            err = ulp_error(got[~m], (np.sin if op == "SIN" else np.cos)(y[~m].astype(np.float64)))
            assert np.all(err <= 2.0), err.max()
        else:
            assert_bits_equal(got, want, f"{op} bits")
    # the reference's own vector-body outputs (golden, valid Cephes domain)
    xin = golden[f"in_{op}_float32"]
    assert_bits_equal(dev.unary(op, dev.to_device(xin)).to_host(), golden[f"ref_{op}_float32"], f"{op} golden")
    # edge cases follow IEEE / NumPy
    e = golden["in_edge_float32"]
    got = dev.unary(op, dev.to_device(e)).to_host()
    err = ulp_error(got, golden[f"np64_{op}_edge"])
    m = np.ones(e.size, bool) if op in ("LOG", "EXP") else ((np.abs(e) <= 3.2) | ~(np.abs(e) <= 8192))
    assert np.all(err[m] <= 2.0), (e[m][err[m] > 2], got[m][err[m] > 2])


@pytest.mark.parametrize("op,lo,hi", [("SIN", -20, 20), ("SIN", -1e5, 1e5), ("COS", -20, 20),
                                      ("LOG", 1e-300, 1e300), ("LOG", 0.5, 2.0), ("LOG", 0.999, 1.001),
                                      ("LOG", 1 - 1e-9, 1 + 1e-9), ("EXP", -700, 700), ("COS", -1e5, 1e5)])
def test_trig_f64_within_2ulp(dev, rng, op, lo, hi):
    if op == "LOG":
        x = np.exp(rng.uniform(np.log(lo), np.log(hi), 1 << 17))
    else:
        x = rng.uniform(lo, hi, 1 << 17)
    x = np.concatenate([x, [0.0, -0.0, np.inf, -np.inf, np.nan, 1.0, -1.0, 5e-324, 1e-310]])
    fn = {"SIN": np.sin, "COS": np.cos, "LOG": np.log, "EXP": np.exp}[op]
    with np.errstate(all="ignore"):
        truth = fn(x.astype(np.longdouble))
    got = dev.unary(op, dev.to_device(x)).to_host()
    with np.errstate(all="ignore"):
        g = got.astype(np.longdouble)
        fin = np.isfinite(truth) & np.isfinite(g)
        sp = np.maximum(np.spacing(np.abs(truth[fin]).astype(np.float64)), 5e-324)
        err = np.abs(g[fin] - truth[fin]) / sp
    assert err.max() <= 2.0, err.max()
    assert np.array_equal(np.isnan(got[~fin]), np.isnan(truth[~fin].astype(np.float64)))
    nn = ~fin & ~np.isnan(got)
    assert np.array_equal(got[nn], truth[nn].astype(np.float64))


@pytest.mark.parametrize("dt", FP_DTYPES)
@pytest.mark.parametrize("op", ["SIN", "COS", "LOG", "EXP"])
def test_trig_results_do_not_depend_on_neighbours(dev, rng, op, dt):
    """The kernels test a thread's whole set of inputs once and take a branch-free fast body when all
    are in its domain, else a general per-element function.  A value must map to the same bits either
    way: sprinkle out-of-domain values (NaN, inf, 0, negatives, huge, denormals) through the array so
    that the same in-domain values are seen once in all-fast packs and once in mixed ones, and compare
    with the all-fast run, the strided kernel and (edge values) NumPy."""
    from pnumpy_b200 import cabi
    n = (1 << 16) + 5
    x = (np.exp(rng.uniform(-20, 20, n)) if op == "LOG" else rng.uniform(-50, 50, n)).astype(dt)
    clean = dev.unary(op, dev.to_device(x)).to_host()
    edges = np.array([np.nan, np.inf, -np.inf, 0.0, -0.0, -1.0, 1e30, -1e30, np.finfo(dt).tiny / 4, 1e6, 2e5, 800.0, -800.0], dtype=dt)
    y = x.copy()
    pos = rng.choice(n, 400, replace=False)
    y[pos] = edges[np.arange(pos.size) % edges.size]
    mixed = dev.unary(op, dev.to_device(y)).to_host()
    keep = np.ones(n, bool)
    keep[pos] = False
    assert np.array_equal(mixed[keep].view(np.uint8), clean[keep].view(np.uint8))
    # edge values follow NumPy (the reference's own edge behaviour is defective)
    fn = {"SIN": np.sin, "COS": np.cos, "LOG": np.log, "EXP": np.exp}[op]
    with np.errstate(all="ignore"):
        want = fn(y[pos].astype(np.longdouble)).astype(dt)
    got = mixed[pos]
    assert np.array_equal(np.isnan(got), np.isnan(want))
    fin = np.isfinite(want)
    assert np.array_equal(got[~fin & ~np.isnan(want)], want[~fin & ~np.isnan(want)])
    with np.errstate(all="ignore"):
        assert np.all(np.abs(got[fin].astype(np.longdouble) - want[fin]) <= 4 * np.maximum(np.spacing(np.abs(want[fin])), np.finfo(dt).smallest_subnormal))
    # the generic strided kernel (one element per lane) gives the same bits as the vector kernel
    wide = np.zeros(2 * n, dtype=dt)
    wide[::2] = y
    dw, do = dev.to_device(wide), dev.empty(n, dt)
    fnp = cabi.get_trig(cabi.TRIG_OPS[op], do.atype)[0]
    cabi.check(fnp(dw.ptr, do.ptr, n, 2 * y.itemsize, y.itemsize, None), "strided")
    dev.sync()
    strided = do.to_host()
    assert np.array_equal(np.isnan(strided), np.isnan(mixed))
    ok = ~np.isnan(mixed)
    assert np.array_equal(strided[ok].view(np.uint8), mixed[ok].view(np.uint8))


# ---- reductions ---------------------------------------------------------------------------------
def _sum_tol(a):
    s = np.abs(a.astype(np.float64)).sum()
    if a.dtype == np.float32:
        return lambda ref: 2.0 ** -23 * abs(ref) + 2.0 ** -50 * s
    return lambda ref: (math.log2(max(a.size, 2)) + 2) * 2.0 ** -53 * s


@pytest.mark.parametrize("dt", DTYPES)
def test_reductions_vs_oracle_and_golden(dev, O, golden, dt):
    a = reduction_input(dt)
    assert zlib.crc32(a.tobytes()) == int(golden[f"red_crc_{dt}"])
    da = dev.to_device(a)
    zero = dev.to_device(np.zeros(1, a.dtype))
    with np.errstate(all="ignore"):
        got = dev.reduce("ADD", da, start=zero).to_host()[0]
        if dt in FP_DTYPES:
            ref = math.fsum(a.astype(np.float64))
            assert abs(float(got) - ref) <= _sum_tol(a)(ref), (got, ref)
            assert abs(float(got) - float(O.reduce("ADD", a, start=0))) <= 2 * _sum_tol(a)(ref)
        else:
            assert got == golden[f"red_SUM_{dt}"] == O.reduce("ADD", a, start=np.zeros(1, a.dtype)[0])
        # NumPy's calling convention for min/max: accumulator pre-loaded with a[0], n-1 elements
        first = dev.to_device(a[:1])
        for op in ("MIN", "MAX"):
            got = dev.reduce(op, da.view(1, a.size - 1), start=first).to_host()[0]
            assert got == golden[f"red_{op}_{dt}"] == O.reduce(op, a[1:], start=a[0]), (op, got)
            assert dev.reduce(op, da).to_host()[0] == golden[f"red_{op}_{dt}"]  # NULL start
        if dt != "bool":
            got = dev.reduce("MUL", da.view(0, 1000), start=dev.to_device(np.ones(1, a.dtype))).to_host()[0]
            want = O.reduce("MUL", a[:1000], start=np.ones(1, a.dtype)[0])
            if dt in FP_DTYPES:
                assert (np.isnan(got) and np.isnan(want)) or got == want or abs(got - want) <= 1e-5 * abs(want)
            else:
                assert got == want
        assert dev.argminmax(0, da).to_host()[0] == int(golden[f"red_ARGMAX_{dt}"]) == O.argminmax(0, a)
        assert dev.argminmax(1, da).to_host()[0] == int(golden[f"red_ARGMIN_{dt}"]) == O.argminmax(1, a)


@pytest.mark.parametrize("dt", FP_DTYPES)
@pytest.mark.parametrize("tag", ["none", "first", "last", "mid"])
def test_nan_reductions(dev, golden, dt, tag):
    a = nan_case_input(dt, tag)
    assert zlib.crc32(a.tobytes()) == int(golden[f"nan_crc_{tag}_{dt}"])
    da = dev.to_device(a)
    for op in ("MIN", "MAX"):
        got = dev.reduce(op, da).to_host()[0]
        want = golden[f"nan_{op}_{tag}_{dt}"]
        assert (np.isnan(got) and np.isnan(want)) or got == want, (op, got, want)
    assert dev.argminmax(0, da).to_host()[0] == int(golden[f"nan_ARGMAX_{tag}_{dt}"])
    assert dev.argminmax(1, da).to_host()[0] == int(golden[f"nan_ARGMIN_{tag}_{dt}"])


@pytest.mark.parametrize("dt", ["int16", "float32", "float64"])
def test_reduction_edge_shapes(dev, O, rng, dt):
    """empty, 1 element, ragged sizes around the 65536-element logical block, strided input,
    first-occurrence ties, run-to-run determinism."""
    from pnumpy_b200 import cabi
    for n in (1, 2, 7, 255, 257, 65535, 65536, 65537, 131072 + 5):
        a = reduction_input(dt, n=n, seed=n)
        da = dev.to_device(a)
        with np.errstate(all="ignore"):
            got = dev.reduce("ADD", da).to_host()[0]
            if dt == "int16":
                assert got == np.add.reduce(a, dtype=a.dtype)
            else:
                ref = math.fsum(a.astype(np.float64))
                assert abs(float(got) - ref) <= _sum_tol(a)(ref) + 1e-300
            assert dev.reduce("MAX", da).to_host()[0] == a.max()
            assert dev.argminmax(0, da).to_host()[0] == a.argmax()
            assert dev.argminmax(1, da).to_host()[0] == a.argmin()
    # empty with a start value: out = start
    st = dev.to_device(np.array([42], dtype=dt))
    assert dev.reduce("ADD", dev.empty(0, dt), start=st).to_host()[0] == 42
    # ties: the extreme value repeated across block boundaries -> smallest index
    a = np.zeros(200000, dtype=dt)
    a[[70000, 65535, 65536, 199999]] = 9
    a[[3, 150000]] = -9
    da = dev.to_device(a)
    assert dev.argminmax(0, da).to_host()[0] == 65535
    assert dev.argminmax(1, da).to_host()[0] == 3
    # strided input (every other element) through the raw launcher
    a = reduction_input(dt, n=100001, seed=5)
    da, out = dev.to_device(a), dev.empty(1, dt)
    fn = cabi.get_reduce(cabi.BINARY_OPS["MAX"], cabi.atop_of(dt))
    cabi.check(fn(da.ptr, out.ptr, None, 50000, 2 * a.dtype.itemsize, None), "strided reduce")
    assert out.to_host()[0] == a[::2][:50000].max()
    # bit-reproducible run to run
    b = reduction_input(dt, n=(1 << 20) + 3, seed=77)
    db = dev.to_device(b)
    r = [dev.reduce("ADD", db).to_host().tobytes() for _ in range(5)]
    assert len(set(r)) == 1


def test_two_pass_building_blocks_are_shard_invariant(dev, rng):
    """pncu_reduce_partials over shards cut at multiples of pncu_reduce_block_elems() + one
    pncu_reduce_finish over the gathered partials == the single-call result, bit for bit
    (the multi-GPU combine of the nte dispatcher and of bench.py --gpus N)."""
    from pnumpy_b200 import cabi
    lib = cabi.load()
    A = lib.pncu_reduce_block_elems()
    n = 11 * A + 1234
    for dt, op in (("float32", "ADD"), ("float64", "ADD"), ("float32", "MIN"), ("int32", "ADD"), ("int8", "MAX")):
        a = reduction_input(dt, n=n, seed=11)
        if op == "MIN":
            a[5 * A + 3] = np.nan
        da = dev.to_device(a)
        whole = dev.reduce(op, da).to_host()
        t, f = cabi.atop_of(dt), cabi.BINARY_OPS[op]
        pb = lib.pncu_reduce_partial_bytes(f, t)
        nb = lib.pncu_reduce_partial_count(t, n)
        assert nb == -(-n * a.dtype.itemsize // 65536)
        parts = dev.empty(nb * pb, np.uint8)
        units = -(-n // A)  # shardable units of A elements
        for shards in (1, 2, 4, 8):
            cabi.check(lib.pncu_memset(parts.ptr, 0xFF, nb * pb, None), "memset")
            per = -(-units // shards)
            for s_ in range(shards):
                e0, e1 = min(n, s_ * per * A), min(n, (s_ + 1) * per * A)
                if e0 >= e1:
                    continue
                slot0 = lib.pncu_reduce_partial_count(t, e0)
                cabi.check(lib.pncu_reduce_partials(f, t, da.ptr + e0 * a.dtype.itemsize, e1 - e0,
                                                    parts.ptr + slot0 * pb, None), "partials")
            out = dev.empty(1, dt)
            cabi.check(lib.pncu_reduce_finish(f, t, parts.ptr, nb, out.ptr, None, None), "finish")
            assert out.to_host().tobytes() == whole.tobytes(), (dt, op, shards)
    # arg-reduction with global index bases
    a = reduction_input("float32", n=n, seed=12)
    a[[7 * A + 5, 2 * A + 1]] = a.max() + 1
    da = dev.to_device(a)
    nb = lib.pncu_reduce_partial_count(12, n)
    parts = dev.empty(nb * 16, np.uint8)
    units = -(-n // A)
    for shards in (1, 3, 8):
        per = -(-units // shards)
        for s_ in range(shards):
            e0, e1 = min(n, s_ * per * A), min(n, (s_ + 1) * per * A)
            if e0 >= e1:
                continue
            slot0 = lib.pncu_reduce_partial_count(12, e0)
            cabi.check(lib.pncu_argminmax_partials(0, 12, da.ptr + e0 * 4, e1 - e0, e0,
                                                   parts.ptr + slot0 * 16, None), "arg partials")
        out = dev.empty(1, np.int64)
        cabi.check(lib.pncu_argminmax_finish(0, 12, parts.ptr, nb, out.ptr, None), "arg finish")
        assert out.to_host()[0] == 2 * A + 1 == a.argmax()


# ---- BASELINE.json sizes: size-independent properties ----------------------------------------------
@pytest.mark.parametrize("dt", ["int32", "int64", "float32", "float64"])
def test_full_size_properties(dev, dt):
    """2^28 elements (config 2/3): structured inputs whose results are known in closed form, so no
    CPU pass over 2^28 elements is needed: a = i mod 253 + 1 (the reference's benchmark data,
    src/pnumpy/benchmark.py:56-66), generated on the device from small pieces."""
    n = 1 << 28
    period = 253 * 4096  # multiple of 253; host builds one period, device tiles it by d2d copies
    host = (np.arange(period, dtype=np.int64) % 253 + 1).astype(dt)
    from pnumpy_b200 import cabi
    lib = cabi.load()
    da = dev.empty(n, dt)
    isz = host.dtype.itemsize
    dpiece = dev.to_device(host)
    off = 0
    while off < n:
        m = min(period, n - off)
        cabi.check(lib.pncu_memcpy_d2d(da.ptr + off * isz, dpiece.ptr, m * isz, None), "tile")
        off += m
    dev.sync()
    # a + a == 2a, a * a, a > a == 0, a / a == 1: check by reductions on the device
    two = dev.binary("ADD", da, da)
    full, rem = divmod(n, 253)
    s_period = sum(range(1, 254))
    exact_sum = full * s_period + sum(range(1, rem + 1))
    if dt in ("int32", "int64"):
        got = int(dev.reduce("ADD", two).to_host()[0])
        want = (2 * exact_sum) % (1 << (8 * isz))
        assert got % (1 << (8 * isz)) == want
    else:
        got = float(dev.reduce("ADD", two).to_host()[0])
        assert abs(got - 2 * exact_sum) <= 2.0 ** -22 * 2 * exact_sum
    assert dev.reduce("MAX", two).to_host()[0] == 506 and dev.reduce("MIN", two).to_host()[0] == 2
    gt = dev.compare("GT", two, da)            # 2a > a everywhere
    assert dev.reduce("MIN", gt).to_host()[0] == 1
    eq = dev.compare("GT", da, da)             # a > a nowhere
    assert dev.reduce("MAX", eq).to_host()[0] == 0
    sq = dev.binary("MUL", da, da)
    assert dev.reduce("MAX", sq).to_host()[0] == 253 * 253
    assert dev.argminmax(0, sq).to_host()[0] == 252     # first a == 253
    assert dev.argminmax(1, sq).to_host()[0] == 0
    if dt in ("float32", "float64"):
        one = dev.binary("DIV", da, da)
        assert dev.reduce("MIN", one).to_host()[0] == 1 and dev.reduce("MAX", one).to_host()[0] == 1
        r = dev.unary("SQRT", sq)                        # sqrt(a*a) == a exactly
        d = dev.binary("SUB", r, da)
        assert dev.reduce("MAX", d).to_host()[0] == 0 and dev.reduce("MIN", d).to_host()[0] == 0
    else:
        q = dev.int_true_divide(two, da)
        assert dev.reduce("MIN", q).to_host()[0] == 2.0 and dev.reduce("MAX", q).to_host()[0] == 2.0


@pytest.mark.parametrize("dt", ["float32", "int64"])
def test_full_size_against_the_oracle(dev, O, dt):
    """BASELINE.json configs[1], [2] and [3] at their stated size, 2^28 elements, compared with the CPU ORACLE on
    the same seeded random inputs -- not with the GPU's own reductions: every element of add / multiply / divide /
    greater (and float32 sqrt / sin / log / exp) must have the oracle's bits, min / max / argmin / argmax (with a NaN
    planted for float32) must be the oracle's, and the sum must be within the stated tolerance of a float64 sum."""
    n = 1 << 28
    rng = np.random.default_rng(1234)
    if dt == "float32":
        a = rng.random(n, dtype=np.float32) * np.float32(253.0) + np.float32(1.0)
        b = rng.random(n, dtype=np.float32) * np.float32(253.0) + np.float32(1.0)
    else:
        a = rng.integers(1, 254, n, dtype=np.int64)
        b = rng.integers(1, 254, n, dtype=np.int64)
    da, db = dev.to_device(a), dev.to_device(b)
    do = dev.empty(n, dt)
    for op in ("ADD", "MUL") + (("DIV",) if dt == "float32" else ()):
        dev.binary(op, da, db, out=do)
        assert_bits_equal(do.to_host(), O.binary(op, a, b), f"{op} {dt} at 2^28")
    if dt == "int64":
        dq = dev.int_true_divide(da, db)
        assert_bits_equal(dq.to_host(), O.int_true_divide(a, b), "int64 true_divide at 2^28")
        dq.free()
    dg = dev.compare("GT", da, db)
    assert_bits_equal(dg.to_host(), O.compare("GT", a, b), f"greater {dt} at 2^28")
    dg.free()
    if dt == "float32":
        for op, src in (("SQRT", a), ("SIN", None), ("LOG", None), ("EXP", None)):
            if src is None:   # the input ranges of SURVEY.md 8d
                lo, hi = {"SIN": (-8192.0, 8192.0), "LOG": (1e-6, 1e6), "EXP": (-80.0, 80.0)}[op]
                src = (rng.random(n, dtype=np.float32) * np.float32(hi - lo) + np.float32(lo)).astype(np.float32)
                db.copy_from_host(src)
                dev.unary(op, db, out=do)
            else:
                dev.unary(op, da, out=do)
            assert_bits_equal(do.to_host(), O.unary(op, src), f"{op} float32 at 2^28")
    # reductions: the oracle's value for min / max / arg*, the float64 reference for the sum
    if dt == "float32":
        s64 = float(np.sum(a, dtype=np.float64))
        got = float(dev.reduce("ADD", da).to_host()[0])
        assert abs(got - s64) <= 2.0 ** -23 * abs(s64) + 2.0 ** -50 * float(np.sum(np.abs(a), dtype=np.float64))
    else:
        assert int(dev.reduce("ADD", da).to_host()[0]) == int(O.reduce("ADD", a))
    for kind in (0, 1):
        assert int(dev.argminmax(kind, da).to_host()[0]) == O.argminmax(kind, a)
    assert dev.reduce("MIN", da).to_host()[0] == O.reduce("MIN", a) == a.min()
    assert dev.reduce("MAX", da).to_host()[0] == O.reduce("MAX", a) == a.max()
    if dt == "float32":
        a[123456789] = np.nan
        da.copy_from_host(a)
        assert np.isnan(dev.reduce("MIN", da).to_host()[0]) and np.isnan(dev.reduce("MAX", da).to_host()[0])
        assert int(dev.argminmax(0, da).to_host()[0]) == O.argminmax(0, a) == 123456789
        assert int(dev.argminmax(1, da).to_host()[0]) == O.argminmax(1, a) == 123456789
    for x in (da, db, do):
        x.free()


# ---- fused chain a*b + c and sum(a*b + c) -------------------------------------------
FUSED_DTYPES = ["int32", "uint32", "int64", "uint64", "float32", "float64"]


@pytest.mark.parametrize("dt", FUSED_DTYPES)
def test_muladd_vs_oracle(dev, O, rng, dt):
    """out = a*b + c in one kernel == the oracle's MUL body then its ADD body (two roundings, no FMA),
    on raw bit patterns, for aligned, misaligned and in-place operands."""
    launches0 = _launches()
    n = N_SMALL
    a, b, c = (fill_random(n + 8, dt, rng) for _ in range(3))
    with np.errstate(all="ignore"):
        want = O.binary("ADD", O.binary("MUL", a, b), c)
        assert_bits_equal(want, a * b + c, f"oracle vs numpy muladd {dt}")
    da, db, dc = dev.to_device(a), dev.to_device(b), dev.to_device(c)
    got = dev.muladd(da, db, dc).to_host()
    assert_bits_equal(got, want, f"muladd {dt}")
    # operands whose 32-byte alignments disagree (element kernel) and agree-but-offset (head/tail path)
    for oa, ob, oc in ((1, 2, 3), (3, 3, 3)):
        out = dev.empty(n, dt)
        dev.muladd(da.view(oa, n), db.view(ob, n), dc.view(oc, n), out=out)
        with np.errstate(all="ignore"):
            w = O.binary("ADD", O.binary("MUL", a[oa:oa + n], b[ob:ob + n]), c[oc:oc + n])
        assert_bits_equal(out.to_host(), w, f"muladd {dt} offsets {(oa, ob, oc)}")
    dev.muladd(da, db, dc, out=dc)  # in place over the addend
    assert_bits_equal(dc.to_host(), want, f"muladd {dt} in place")
    assert _launches() - launches0 >= 4


@pytest.mark.parametrize("dt", FUSED_DTYPES)
def test_muladd_sum_is_the_three_call_chain(dev, O, dt):
    """sum(a*b + c) fused == multiply, add, sum as three kernels, BIT FOR BIT (same logical blocks, same
    fold order, same accumulator), and within the stated sum tolerance of a float64 reference."""
    from pnumpy_b200 import cabi
    A = cabi.load().pncu_reduce_block_elems()
    for n in (1, 777, A // 4 + 5, 3 * A + 1234):
        a, b, c = (reduction_input(dt, n=n, seed=s) for s in (21, 22, 23))
        da, db, dc = dev.to_device(a), dev.to_device(b), dev.to_device(c)
        chain = dev.reduce("ADD", dev.binary("ADD", dev.binary("MUL", da, db), dc)).to_host()
        fused = dev.muladd_sum(da, db, dc).to_host()
        assert fused.tobytes() == chain.tobytes(), (dt, n, fused, chain)
        with np.errstate(all="ignore"):
            u = a * b + c
        if dt in FP_DTYPES:
            ref = math.fsum(u.astype(np.float64))
            mag = float(np.abs(u.astype(np.float64)).sum())
            tol = (2.0 ** -23 * abs(ref) + 2.0 ** -50 * mag) if dt == "float32" else (math.log2(max(n, 2)) + 2) * 2.0 ** -53 * mag
            assert abs(float(fused[0]) - ref) <= tol, (dt, n, fused, ref)
        else:
            assert fused[0] == u.sum(dtype=u.dtype), (dt, n)


def test_muladd_sum_partials_are_shard_invariant(dev):
    """pass 1 of the fused sum over shards + the ordinary ADD finish == the single call (multi-GPU combine)."""
    from pnumpy_b200 import cabi
    lib = cabi.load()
    A = lib.pncu_reduce_block_elems()
    n = 9 * A + 4321
    for dt in ("float32", "float64", "int64"):
        a, b, c = (reduction_input(dt, n=n, seed=s) for s in (31, 32, 33))
        da, db, dc = dev.to_device(a), dev.to_device(b), dev.to_device(c)
        whole = dev.muladd_sum(da, db, dc).to_host()
        t = cabi.atop_of(dt)
        pb = lib.pncu_reduce_partial_bytes(cabi.BINARY_OPS["ADD"], t)
        nb = lib.pncu_reduce_partial_count(t, n)
        parts = dev.empty(nb * pb, np.uint8)
        units, isz = -(-n // A), a.dtype.itemsize
        for shards in (1, 2, 4, 8):
            cabi.check(lib.pncu_memset(parts.ptr, 0xFF, nb * pb, None), "memset")
            per = -(-units // shards)
            for s_ in range(shards):
                e0, e1 = min(n, s_ * per * A), min(n, (s_ + 1) * per * A)
                if e0 >= e1:
                    continue
                slot0 = lib.pncu_reduce_partial_count(t, e0)
                cabi.check(lib.pncu_muladd_sum_partials(t, da.ptr + e0 * isz, db.ptr + e0 * isz, dc.ptr + e0 * isz,
                                                        e1 - e0, parts.ptr + slot0 * pb, None), "muladd partials")
            out = dev.empty(1, dt)
            cabi.check(lib.pncu_reduce_finish(cabi.BINARY_OPS["ADD"], t, parts.ptr, nb, out.ptr, None, None), "finish")
            assert out.to_host().tobytes() == whole.tobytes(), (dt, shards)


def test_muladd_full_size_properties(dev):
    """2^28 float32 elements: a*1 + 0 reproduces a bit for bit (no NaNs in the input), sum(a*1 + 0) is
    sum(a) bit for bit, and sum(a*b + c) == sum over the materialised a*b + c."""
    n = 1 << 28
    rng = np.random.default_rng(99)
    a = rng.random(n, dtype=np.float32) * 253 + 1
    da = dev.to_device(a)
    one = dev.to_device(np.ones(n, np.float32))
    zero = dev.to_device(np.zeros(n, np.float32))
    out = dev.muladd(da, one, zero)
    eq = dev.compare("EQ", out, da)
    assert bool(dev.reduce("MUL", eq).to_host()[0])  # logical AND over the bool result
    assert dev.muladd_sum(da, one, zero).to_host().tobytes() == dev.reduce("ADD", da).to_host().tobytes()
    del eq, zero
    u = dev.muladd(da, da, one, out=out)
    assert dev.muladd_sum(da, da, one).to_host().tobytes() == dev.reduce("ADD", u).to_host().tobytes()


# ---- the sixteen transcendental ufuncs the reference hooks without a body ------------
# (numpy function, input range) per op; the bar is <= 4 ULP of the correctly rounded value (CUDA documents
# 1-4 ULP for these; NumPy's own loops differ between libm builds in the last place, so there are no
# reference bits to match) and exact IEEE special cases.
MORE_TRIG = {
    "TAN": (np.tan, -20.0, 20.0), "ASIN": (np.arcsin, -1.0, 1.0), "ACOS": (np.arccos, -1.0, 1.0),
    "ATAN": (np.arctan, -50.0, 50.0), "SINH": (np.sinh, -20.0, 20.0), "COSH": (np.cosh, -20.0, 20.0),
    "TANH": (np.tanh, -10.0, 10.0), "ASINH": (np.arcsinh, -1e4, 1e4), "ACOSH": (np.arccosh, 1.0, 1e4),
    "ATANH": (np.arctanh, -1.0, 1.0), "LOG2": (np.log2, 1e-30, 1e30), "LOG10": (np.log10, 1e-30, 1e30),
    "EXP2": (np.exp2, -100.0, 100.0), "EXPM1": (np.expm1, -20.0, 20.0), "LOG1P": (np.log1p, -0.999, 1e4),
    "CBRT": (np.cbrt, -1e6, 1e6),
}


@pytest.mark.parametrize("dt", FP_DTYPES)
@pytest.mark.parametrize("op", sorted(MORE_TRIG))
def test_more_transcendentals_within_4ulp(dev, O, rng, op, dt):
    fn, lo, hi = MORE_TRIG[op]
    n = 1 << 16
    if op in ("LOG2", "LOG10"):
        x = np.exp(rng.uniform(np.log(lo), np.log(hi), n))
    else:
        x = rng.uniform(lo, hi, n)
    small = rng.uniform(-1e-3, 1e-3, 1 << 12)                      # cancellation region of expm1/log1p/atanh/...
    if op == "ACOSH":
        small = 1.0 + np.abs(small)
    elif op in ("LOG2", "LOG10"):
        small = 1.0 + small
    # magnitudes from 1e-30 up (log-uniform): where asinh/atanh/sinh/... are x (1 + O(x^2)) and a sloppy reciprocal
    # or a cancelling sum shows; acosh gets 1 + tiny, the logs 1 +- tiny
    tiny = np.exp(rng.uniform(np.log(1e-30), 0.0, 1 << 12)) * rng.choice([-1.0, 1.0], 1 << 12)
    if op == "ACOSH":
        tiny = 1.0 + np.abs(tiny)
    elif op in ("LOG2", "LOG10"):
        tiny = 1.0 + tiny * 0.5
    elif op == "LOG1P":
        tiny = np.where(tiny < -0.5, -tiny, tiny)
    edge = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1.0, -1.0, 2.0, -2.0, 0.5, 1e-40, -1e-40, 1e38, -1e38,
                     1.0 - 2.0 ** -24, -1.0 + 2.0 ** -24, 1.0 + 2.0 ** -23, 2.0 ** 59, 2.0 ** 60, 2.0 ** 61, -2.0 ** 60])
    x = np.concatenate([x, small, tiny, edge]).astype(dt)
    launches0 = _launches()
    got = dev.unary(op, dev.to_device(x)).to_host()
    assert _launches() > launches0 and got.dtype == x.dtype
    with np.errstate(all="ignore"):
        truth = fn(x.astype(np.longdouble))                        # x87 extended: 11 spare bits for float64
        want = O.unary(op, x)                                      # glibc, what NumPy's loop calls
    tf = truth.astype(np.float64)
    g = got.astype(np.longdouble)
    fin = np.isfinite(truth) & np.isfinite(got)
    with np.errstate(all="ignore"):
        tiny = float(np.finfo(x.dtype).smallest_subnormal)
        sp = np.maximum(np.spacing(np.abs(truth[fin]).astype(x.dtype)).astype(np.float64), tiny)
        err = np.abs(g[fin] - truth[fin]).astype(np.float64) / sp
    assert err.max() <= 4.0, (op, dt, err.max(), x[fin][np.argmax(err)])
    # non-finite results: same class as the libm oracle (NaN <-> NaN, +-inf exact), also where only one
    # side overflowed the dtype
    rest = ~fin
    assert np.array_equal(np.isnan(got[rest]), np.isnan(want[rest])), (op, dt, x[rest], got[rest], want[rest])
    nn = rest & ~np.isnan(got)
    assert np.array_equal(got[nn], want[nn]), (op, dt, x[nn], got[nn], want[nn])
    # and never far from the libm oracle either
    both = np.isfinite(want) & np.isfinite(got)
    assert np.all(ulp_error(got[both], want[both].astype(np.float64)) <= 8.0)
    del tf


NEIGHBOUR_CASES = [(op, "float64") for op in sorted(MORE_TRIG)] + [(op, "float32") for op in ("ASINH", "ACOSH", "ATANH")]


@pytest.mark.parametrize("op,dt", NEIGHBOUR_CASES)
def test_more_transcendentals_do_not_depend_on_neighbours(dev, rng, op, dt):
    """float64 (all) and the hand-written float32 bodies: the lean body runs when a thread's whole set of inputs
    is in the fast domain, the general function otherwise -- a value must get the same bits either way, and from
    the strided kernel."""
    from pnumpy_b200 import cabi
    fn, lo, hi = MORE_TRIG[op]
    n = (1 << 15) + 3
    isz = np.dtype(dt).itemsize
    x = rng.uniform(lo, hi, n).astype(dt)
    clean = dev.unary(op, dev.to_device(x)).to_host()
    y = x.copy()
    pos = rng.choice(n, 300, replace=False)
    big = 1e300 if dt == "float64" else 1e38
    edges = np.array([np.nan, np.inf, -np.inf, big, -big, 0.0, -0.0, 1.0, -1.0, float(np.finfo(dt).smallest_subnormal),
                      2.0 ** 80, -3.0]).astype(dt)
    y[pos] = edges[np.arange(pos.size) % edges.size]
    mixed = dev.unary(op, dev.to_device(y)).to_host()
    keep = np.ones(n, bool)
    keep[pos] = False
    assert np.array_equal(mixed[keep].view(np.uint8), clean[keep].view(np.uint8)), op
    wide = np.zeros(2 * n, dt)
    wide[::2] = y
    dw, do = dev.to_device(wide), dev.empty(n, dt)
    fnp = cabi.get_trig(cabi.TRIG_OPS[op], do.atype)[0]
    cabi.check(fnp(dw.ptr, do.ptr, n, 2 * isz, isz, None), "strided")
    dev.sync()
    strided = do.to_host()
    assert np.array_equal(np.isnan(strided), np.isnan(mixed))
    ok = ~np.isnan(mixed)
    assert np.array_equal(strided[ok].view(np.uint8), mixed[ok].view(np.uint8)), op


# ---- the cross-GPU exchange fused into the two passes -------------------------------------------------
def _exchange_case(dev, devices, rng, group_aligned=False, groups_per_shard=2):
    """sum / min / max / argmax of one array cut into len(devices) shards, every shard on its own GPU, through
    Exchange: every participant must end with the bits a single GPU produces.  group_aligned: the shards start at
    multiples of the fold group (64 slots), so each participant folds its own groups and only group partials travel
    (pncu_fused_exchange::groups) -- same bits again."""
    import ctypes
    from pnumpy_b200 import cabi, exchange
    lib = cabi.load()
    world = len(devices)
    timeouts0 = {}
    align = int(lib.pncu_reduce_block_elems())
    n = 5 * align * world + 12345
    for dt in ("float32", "float64", "int32"):
        if group_aligned:
            align = int(lib.pncu_reduce_group_slots()) * (65536 // np.dtype(dt).itemsize)
            n = groups_per_shard * align * world + 12345
        a = (rng.normal(0, 100, n)).astype(dt)
        units = -(-n // align)
        per = -(-units // world)
        bounds = [min(n, r * per * align) for r in range(world + 1)]
        if dt != "int32":
            a[n // 2 + 7] = np.nan
        else:
            # SURVEY.md 8d case (v): the extreme value tied on both sides of a shard (or block) boundary --
            # the first occurrence must win whichever participant saw it
            edge = bounds[1] if world > 1 else 3 * align
            a[edge - 1] = a[edge] = a[edge + 5] = 10 ** 6
            a[edge - 2] = a[edge + 1] = -(10 ** 6)
        cabi.check(lib.pncu_set_device(devices[0]), "set_device")
        whole = dev.to_device(a)
        want = {op: dev.reduce(op, whole).to_host().tobytes() for op in ("ADD", "MIN", "MAX")}
        want_arg = [int(dev.argminmax(k, whole).to_host()[0]) for k in (0, 1)]
        xs = exchange.Exchange.in_process(devices, max_slots=max(8192, 2 * groups_per_shard * 64 * world + 64))
        shards, outs, idxs, streams = [], [], [], []
        for r, d in enumerate(devices):
            cabi.check(lib.pncu_set_device(d), "set_device")
            shards.append(dev.to_device(a[bounds[r]:bounds[r + 1]]))
            outs.append(dev.empty(1, dt))
            idxs.append(dev.empty(1, np.int64))
            # participants wait for each other ON THE DEVICE: two that share a GPU need a stream each
            st = ctypes.c_void_p()
            cabi.check(lib.pncu_stream_create(ctypes.byref(st)), "stream")
            streams.append(st)
            timeouts0[d] = int(lib.pncu_exchange_timeouts())
        counts = exchange.block_counts([s.size for s in shards], dt)
        for rep in range(3):  # several calls: both tables, growing counters
            for op in ("ADD", "MIN", "MAX"):
                for r, d in enumerate(devices):
                    cabi.check(lib.pncu_set_device(d), "set_device")
                    xs[r].reduce(op, shards[r], counts, outs[r], stream=streams[r])
                for r, d in enumerate(devices):
                    cabi.check(lib.pncu_set_device(d), "set_device")
                    xs[r].wait(streams[r])
                    got = outs[r].to_host()
                    if op != "ADD" and dt != "int32":
                        assert np.isnan(got[0])
                    else:
                        assert got.tobytes() == want[op], (dt, op, r, rep)
            for kind in (0, 1):
                for r, d in enumerate(devices):
                    cabi.check(lib.pncu_set_device(d), "set_device")
                    xs[r].argminmax(kind, shards[r], counts, bounds[r], idxs[r], stream=streams[r])
                for r, d in enumerate(devices):
                    cabi.check(lib.pncu_set_device(d), "set_device")
                    xs[r].wait(streams[r])
                    assert int(idxs[r].to_host()[0]) == want_arg[kind], (dt, kind, r, rep)
        for r, d in enumerate(devices):
            cabi.check(lib.pncu_set_device(d), "set_device")
            assert lib.pncu_exchange_timeouts() == timeouts0[d]
            assert (xs[r].hierarchical > 0) == (group_aligned or world == 1), (xs[r].hierarchical, group_aligned)
            xs[r].close()
            cabi.check(lib.pncu_stream_destroy(streams[r]), "stream")
    cabi.check(lib.pncu_set_device(devices[0]), "set_device")


def test_fused_exchange_one_gpu(dev, rng):
    """world = 1: the one-launch path (ticket, last CTA folds) equals the plain two passes"""
    _exchange_case(dev, [0], rng)


def test_fused_exchange_two_participants_on_one_gpu(dev, rng):
    """two participants that happen to share GPU 0: the whole exchange protocol (stores into the other
    participant's table, system fences, arrival counters, device-side wait, alternating tables) runs on a
    single-GPU box too"""
    _exchange_case(dev, [0, 0], rng)


def test_fused_exchange_group_partials_on_one_gpu(dev, rng):
    """shards that start at whole fold groups: level A of pass 2 is done by each shard's owner and only the group
    partials are stored into the consumers' tables"""
    _exchange_case(dev, [0, 0], rng, group_aligned=True)


def test_fused_exchange_group_tickets_on_one_gpu(dev, rng):
    """shards of >= 256 slots that start at whole groups: the GROUP-TICKET form of the one-launch kernels (every group is
    folded during pass 1 by the CTA that completes it) together with the exchange of group partials, all-reduce form"""
    _exchange_case(dev, [0, 0], rng, group_aligned=True, groups_per_shard=5)


def test_fused_exchange_group_tickets_two_gpus(dev, rng):
    from pnumpy_b200 import cabi
    if cabi.load().pncu_device_count() < 2:
        pytest.skip("needs two GPUs")
    _exchange_case(dev, [0, 1], rng, group_aligned=True, groups_per_shard=5)


def test_fused_exchange_group_partials_two_gpus(dev, rng):
    from pnumpy_b200 import cabi
    if cabi.load().pncu_device_count() < 2:
        pytest.skip("needs two GPUs")
    _exchange_case(dev, [0, 1], rng, group_aligned=True)


def test_exchange_timeout_is_loud(dev, rng):
    """A participant that never shows up must be an ERROR: the waiting kernel gives up after its bounded wait
    WITHOUT folding the incomplete table and without writing `out`; Exchange.wait() raises and the Exchange
    refuses further use."""
    from pnumpy_b200 import cabi, exchange
    lib = cabi.load()
    n = 4 * int(lib.pncu_reduce_block_elems())
    a = rng.normal(0, 1, n).astype(np.float32)
    xs = exchange.Exchange.in_process([0, 0], max_slots=4096)   # participant 1 never launches anything
    shard = dev.to_device(a)
    out = dev.to_device(np.array([-777.0], dtype=np.float32))
    counts = exchange.block_counts([n, n], np.float32)
    before = int(lib.pncu_exchange_timeouts())
    xs[0].reduce("ADD", shard, counts, out)
    with pytest.raises(cabi.PnumpyCudaError, match="did not deliver"):
        xs[0].wait()
    assert int(lib.pncu_exchange_timeouts()) == before + 1
    assert out.to_host()[0] == np.float32(-777.0)               # never written: no partial fold leaked out
    with pytest.raises(cabi.PnumpyCudaError, match="unusable"):
        xs[0].reduce("ADD", shard, counts, out)
    # the stream's ticket was re-armed by the kernel itself: ordinary reductions keep working
    assert_bits_equal(dev.reduce("ADD", shard).to_host(), dev.reduce("ADD", shard).to_host())
    cabi.check(lib.pncu_fused_reset(None), "fused_reset")
    for x in xs:
        x.close()


def test_one_launch_reductions_match_the_two_pass_building_blocks(dev, rng):
    """pncu_GetReduceMathOpFast / pncu_GetArgMinMaxOpFast launchers are ONE kernel (last CTA folds); the result
    must have the bits of pncu_reduce_partials + pncu_reduce_finish for every type, at sizes that give 1 block,
    a ragged block, fewer and more partials than the 1024 virtual finish threads."""
    from pnumpy_b200 import cabi
    lib = cabi.load()
    for dt in ("int8", "uint16", "int32", "int64", "float32", "float64"):
        isz = np.dtype(dt).itemsize
        be = 65536 // isz
        for n in (1, be - 1, be, 3 * be + 17, 1500 * be + 5, 9000 * be):
            if n * isz > (1 << 30):
                continue
            a = fill_random(n, dt, rng) if np.dtype(dt).kind != "f" else rng.normal(0, 1e3, n).astype(dt)
            d = dev.to_device(a)
            nb = int(lib.pncu_reduce_partial_count(cabi.atop_of(dt), n))
            for op in ("ADD", "MIN", "MAX"):
                f, t = cabi.BINARY_OPS[op], cabi.atop_of(dt)
                parts = dev.empty(nb * lib.pncu_reduce_partial_bytes(f, t), np.uint8)
                two = dev.empty(1, dt)
                cabi.check(lib.pncu_reduce_partials(f, t, d.ptr, n, parts.ptr, None), "partials")
                cabi.check(lib.pncu_reduce_finish(f, t, parts.ptr, nb, two.ptr, None, None), "finish")
                assert dev.reduce(op, d).to_host().tobytes() == two.to_host().tobytes(), (dt, n, op)
            for kind in (0, 1):
                parts = dev.empty(nb * 16, np.uint8)
                two = dev.empty(1, np.int64)
                cabi.check(lib.pncu_argminmax_partials(kind, cabi.atop_of(dt), d.ptr, n, 0, parts.ptr, None), "arg partials")
                cabi.check(lib.pncu_argminmax_finish(kind, cabi.atop_of(dt), parts.ptr, nb, two.ptr, None), "arg finish")
                assert int(dev.argminmax(kind, d).to_host()[0]) == int(two.to_host()[0]), (dt, n, kind)
            d.free()


def test_fused_exchange_two_gpus(dev, rng):
    """two GPUs in one process: pass 1 on each stores into both slot tables over NVLink"""
    from pnumpy_b200 import cabi
    if cabi.load().pncu_device_count() < 2:
        pytest.skip("needs two GPUs")
    _exchange_case(dev, [0, 1], rng)


@pytest.mark.parametrize("dt", ["int8", "uint8", "int16", "uint16", "int32", "float32"])
def test_extreme_in_the_ragged_last_block(dev, dt):
    """Regression: the 1-byte word-at-a-time min/max of FULL 64 KiB blocks produced packed partials that compared
    wrongly with the plain partial of a ragged last block (signed bytes were not sign-extended).  Extremes of
    either sign, in the tail and in the body, several block counts."""
    be = 65536 // np.dtype(dt).itemsize
    signed = np.dtype(dt).kind != "u"
    for nblocks in (1, 2, 5):
        for tail in (1, 7, 300):
            n = nblocks * be + tail
            for body, tailv in ((1, 100), (100, 1), (-3 if signed else 2, -100 if signed else 0), (-100 if signed else 0, -3 if signed else 2)):
                a = np.full(n, 50, dtype=dt)
                a[be // 2] = body
                a[n - 1 - tail // 2] = tailv
                d = dev.to_device(a)
                assert dev.reduce("MAX", d).to_host()[0] == a.max(), (dt, n, body, tailv)
                assert dev.reduce("MIN", d).to_host()[0] == a.min(), (dt, n, body, tailv)
                assert int(dev.argminmax(0, d).to_host()[0]) == int(a.argmax())
                assert int(dev.argminmax(1, d).to_host()[0]) == int(a.argmin())


def test_reductions_at_2p30_with_nan_and_64bit_indices(dev):
    """configs[3] at its stated size on one GPU: 2^30 float32 (4 GiB) with NaNs planted, and a 2^32+5-element
    int8 array whose extreme sits beyond the 32-bit index range."""
    from pnumpy_b200 import cabi
    lib = cabi.load()

    def tiled(n, dt):
        period = 253 * 4096
        host = (np.arange(period, dtype=np.int64) % 253 + 1).astype(dt)
        d = dev.empty(n, dt)
        piece = dev.to_device(host)
        off, isz = 0, host.dtype.itemsize
        while off < n:
            m = min(period, n - off)
            cabi.check(lib.pncu_memcpy_d2d(d.ptr + off * isz, piece.ptr, m * isz, None), "tile")
            off += m
        dev.sync()
        piece.free()
        return d

    def poke(d, index, value):
        v = np.array([value], dtype=d.dtype)
        cabi.check(lib.pncu_memcpy_h2d(d.ptr + index * d.dtype.itemsize, v.ctypes.data, d.dtype.itemsize, None), "poke")
        dev.sync()

    n = 1 << 30
    a = tiled(n, np.float32)
    full, rem = divmod(n, 253)
    exact = full * sum(range(1, 254)) + sum(range(1, rem + 1))
    got = float(dev.reduce("ADD", a).to_host()[0])
    assert abs(got - exact) <= 2.0 ** -23 * exact            # float64 accumulation, rounded once
    assert dev.argminmax(0, a).to_host()[0] == 252 and dev.argminmax(1, a).to_host()[0] == 0
    poke(a, n - 1, np.nan)                                     # (iii) NaN in the last element
    assert np.isnan(dev.reduce("MIN", a).to_host()[0]) and np.isnan(dev.reduce("MAX", a).to_host()[0])
    assert np.isnan(dev.reduce("ADD", a).to_host()[0])
    assert dev.argminmax(0, a).to_host()[0] == n - 1 and dev.argminmax(1, a).to_host()[0] == n - 1
    mid = 5 * (1 << 27) + 3
    poke(a, mid, np.nan)                                       # an earlier NaN wins
    assert dev.argminmax(0, a).to_host()[0] == mid and dev.argminmax(1, a).to_host()[0] == mid
    poke(a, 0, np.nan)                                         # (ii) NaN at index 0
    assert dev.argminmax(0, a).to_host()[0] == 0
    a.free()

    m = (1 << 32) + 5                                          # int8: 4 GiB, indices beyond 32 bits
    b = dev.empty(m, np.int8)
    cabi.check(lib.pncu_memset(b.ptr, 1, m, None), "memset")
    dev.sync()
    poke(b, (1 << 32) + 2, 5)
    poke(b, (1 << 32) + 4, -3)
    assert dev.reduce("MAX", b).to_host()[0] == 5 and dev.reduce("MIN", b).to_host()[0] == -3
    assert int(dev.argminmax(0, b).to_host()[0]) == (1 << 32) + 2
    assert int(dev.argminmax(1, b).to_host()[0]) == (1 << 32) + 4
    c = dev.binary("ADD", b, b)                                # element-wise grid over > 2^32 elements
    assert dev.reduce("MAX", c).to_host()[0] == 10 and int(dev.argminmax(0, c).to_host()[0]) == (1 << 32) + 2
    assert int(dev.argminmax(1, c).to_host()[0]) == (1 << 32) + 4 and dev.reduce("MIN", c).to_host()[0] == -6
    b.free()
    c.free()


@pytest.mark.parametrize("dt", DTYPES)
def test_random_shapes_and_alignments(dev, rng, dt):
    """Differential sweep over what the reference's tests do not pin: lengths around the 32-byte vector and the
    64 KiB reduction-block boundaries, device pointers at every element alignment (views into a larger buffer),
    extremes anywhere.  Element-wise ops bit-exact against NumPy, reductions exact (integers, min/max, arg) or
    within the stated tolerance (float sums)."""
    isz = np.dtype(dt).itemsize
    be = 65536 // isz
    big = 3 * be + 300
    if np.dtype(dt).kind == "f":
        host_a = rng.normal(0, 1000, big + 64).astype(dt)
        host_b = (rng.normal(0, 1000, big + 64) + 0.5).astype(dt)
    elif dt == "bool":
        host_a = rng.integers(0, 2, big + 64).astype(dt)
        host_b = rng.integers(0, 2, big + 64).astype(dt)
    else:
        info = np.iinfo(dt)
        host_a = rng.integers(info.min, info.max, big + 64, dtype=dt, endpoint=True)
        host_b = rng.integers(info.min, info.max, big + 64, dtype=dt, endpoint=True)
    da, db = dev.to_device(host_a), dev.to_device(host_b)
    lengths = [1, 2, 31, 32, 33, 255, 257, 1000, be - 1, be, be + 1, be + 7, 2 * be, 2 * be + 300, 3 * be + 255]
    for n in lengths:
        for off in (0, 1, 3, int(rng.integers(4, 60))):
            a, b = host_a[off:off + n], host_b[off + 1:off + 1 + n]
            va, vb = da.view(off, n), db.view(off + 1, n)
            with np.errstate(all="ignore"):
                if dt != "bool":
                    assert_bits_equal(dev.binary("ADD", va, vb).to_host(), a + b, f"add {dt} n={n} off={off}")
                    assert_bits_equal(dev.binary("SUB", va, vb).to_host(), a - b, f"sub {dt} n={n} off={off}")
                    assert_bits_equal(dev.binary("MAX", va, vb).to_host(), np.maximum(a, b), f"maximum {dt} n={n} off={off}")
                    assert_bits_equal(dev.unary("ABS", va).to_host(), np.abs(a), f"abs {dt} n={n} off={off}")
                    # scalar second / first operand (stride 0), and in place on a misaligned view
                    sc = db.view(off + 7, 1)
                    assert_bits_equal(dev.binary("SUB", va, sc, b_scalar=True).to_host(), a - host_b[off + 7], f"a-s {dt} n={n} off={off}")
                    assert_bits_equal(dev.binary("SUB", sc, va, a_scalar=True).to_host(), host_b[off + 7] - a, f"s-a {dt} n={n} off={off}")
                    scratch = dev.to_device(host_a[:off + n + 8])
                    vs = scratch.view(off, n)
                    dev.binary("ADD", vs, vb, out=vs)
                    whole = scratch.to_host()
                    assert_bits_equal(whole[off:off + n], a + b, f"in place {dt} n={n} off={off}")
                    assert_bits_equal(whole[:off], host_a[:off], "bytes before the view untouched")
                    assert_bits_equal(whole[off + n:], host_a[off + n:off + n + 8], "bytes after the view untouched")
                    scratch.free()
                assert_bits_equal(dev.compare("GT", va, vb).to_host(), a > b, f"greater {dt} n={n} off={off}")
                assert_bits_equal(dev.compare("NE", va, vb).to_host(), a != b, f"not_equal {dt} n={n} off={off}")
                assert_bits_equal(dev.compare("LTE", va, sc, b_scalar=True).to_host(), a <= host_b[off + 7], f"a<=s {dt} n={n} off={off}") if dt != "bool" else None
                assert dev.reduce("MAX", va).to_host()[0] == a.max(), (dt, n, off)
                assert dev.reduce("MIN", va).to_host()[0] == a.min(), (dt, n, off)
                assert int(dev.argminmax(0, va).to_host()[0]) == int(a.argmax()), (dt, n, off)
                assert int(dev.argminmax(1, va).to_host()[0]) == int(a.argmin()), (dt, n, off)
                got = dev.reduce("ADD", va).to_host()[0]
                if np.dtype(dt).kind == "f":
                    ref = math.fsum(a.astype(np.float64))
                    assert abs(float(got) - ref) <= _sum_tol(a)(ref), (dt, n, off, got, ref)
                elif dt == "bool":
                    assert bool(got) == bool(a.any())
                else:
                    assert got == np.add.reduce(a, dtype=a.dtype), (dt, n, off)


@pytest.mark.parametrize("dt", DTYPES)
def test_every_op_on_misaligned_views(dev, rng, dt):
    """every element-wise op the getters offer for this dtype, on views that start at odd element offsets and
    end off the vector grid, against NumPy's own loops (bit-exact; float results compared as bits, NaN <-> NaN)"""
    from pnumpy_b200 import cabi
    kind = np.dtype(dt).kind
    n_all = 70000
    if kind == "f":
        ha = rng.normal(0, 50, n_all).astype(dt)
        hb = rng.normal(0, 50, n_all).astype(dt)
        ha[::97] = np.nan; hb[::89] = np.inf; ha[::101] = -0.0; hb[::103] = 0.0
    elif dt == "bool":
        ha = rng.integers(0, 2, n_all).astype(dt); hb = rng.integers(0, 2, n_all).astype(dt)
    else:
        info = np.iinfo(dt)
        ha = rng.integers(info.min, info.max, n_all, dtype=dt, endpoint=True)
        hb = rng.integers(info.min, info.max, n_all, dtype=dt, endpoint=True)
        hb[::7] = rng.integers(0, 8 * np.dtype(dt).itemsize + 3, hb[::7].size).astype(dt)  # shift counts incl. >= bits
    da, db = dev.to_device(ha), dev.to_device(hb)
    binary = {"ADD": np.add, "SUB": np.subtract, "MUL": np.multiply, "MIN": np.minimum, "MAX": np.maximum,
              "BITWISE_AND": np.bitwise_and, "BITWISE_OR": np.bitwise_or, "BITWISE_XOR": np.bitwise_xor,
              "BITWISE_LSHIFT": np.left_shift, "BITWISE_RSHIFT": np.right_shift, "DIV": np.true_divide,
              "FLOORDIV": np.floor_divide, "LOGICAL_AND": np.logical_and, "LOGICAL_OR": np.logical_or}
    unary = {"ABS": np.abs, "NEGATIVE": np.negative, "SIGN": np.sign, "INVERT": np.invert, "LOGICAL_NOT": np.logical_not,
             "ISNAN": np.isnan, "ISINF": np.isinf, "ISFINITE": np.isfinite, "FLOOR": np.floor, "CEIL": np.ceil,
             "TRUNC": np.trunc, "ROUND": np.rint, "SQRT": np.sqrt, "SIGNBIT": np.signbit}
    compares = {"EQ": np.equal, "NE": np.not_equal, "LT": np.less, "GT": np.greater, "LTE": np.less_equal, "GTE": np.greater_equal}
    tested = 0
    for off, n in ((1, 4099), (3, 65537), (5, 31), (0, 69990)):
        a, b = ha[off:off + n], hb[off + 2:off + 2 + n]
        va, vb = da.view(off, n), db.view(off + 2, n)
        with np.errstate(all="ignore"):
            for name, f in binary.items():
                fn, out_t = cabi.get_binary(cabi.BINARY_OPS[name], va.atype)
                if fn is None:
                    continue
                if dt == "bool" and name in ("SUB",):
                    continue
                want = f(a, b)
                got = dev.binary(name, va, vb).to_host()
                if dt == "bool":
                    want = want.astype(np.bool_)
                assert got.dtype == want.dtype, (name, dt, got.dtype, want.dtype)
                assert_bits_equal(got, want, f"{name} {dt} off={off} n={n}")
                tested += 1
            for name, f in unary.items():
                fn, out_t = cabi.get_unary(cabi.UNARY_OPS[name], va.atype)
                if fn is None:
                    continue
                if dt == "bool" and name in ("NEGATIVE", "SIGN", "ABS"):
                    continue
                want = f(a)
                got = dev.unary(name, va).to_host()
                assert got.dtype == want.dtype, (name, dt, got.dtype, want.dtype)
                assert_bits_equal(got, want, f"{name} {dt} off={off} n={n}")
                tested += 1
            for name, f in compares.items():
                assert_bits_equal(dev.compare(name, va, vb).to_host(), f(a, b), f"{name} {dt} off={off} n={n}")
                tested += 1
    assert tested >= 4 * 8


# ---- dtype conversion --------------------------------------------------------------
@pytest.mark.parametrize("di", DTYPES)
def test_convert_vs_oracle_and_numpy(dev, O, rng, di):
    """pncu_GetConvertFast: every (from, to) pair on random bit patterns plus edge values, at a length that is not
    a multiple of any vector width and from a misaligned start: bit-exact against the oracle (= NumPy's astype,
    tests/test_oracle.py::test_oracle_convert_is_numpy_astype)."""
    import warnings
    n = N_SMALL
    a = fill_random(n + 3, di, rng)
    if np.dtype(di).kind == "f":
        a[::3] = rng.normal(0, 1e5, a[::3].size).astype(di)
        edge = np.array([2**31, -2**31, 2**31 - 1, 2**63, -2**63, 2**64, 255.9, -0.9, 65536.5, -32769.0, 3e9, 4294967295.0,
                         4294967296.0, -1.0, 0.0, -0.0, np.inf, -np.inf, np.nan, -2147483649.0, 2147483647.5], dtype=di)
        a[1::7] = edge[np.arange(a[1::7].size) % edge.size]
    da = dev.to_device(a)
    for do in DTYPES:
        for off in (0, 3):
            src = a[off: off + n]
            got = dev.astype(da.view(off, n), do).to_host()
            want = O.astype(src, do)
            assert_bits_equal(got, want, f"{di} -> {do} (offset {off})")
            if not (np.dtype(di).kind == "f" and do == "uint32"):
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    assert_bits_equal(got, src.astype(do), f"{di} -> {do} vs NumPy (offset {off})")
