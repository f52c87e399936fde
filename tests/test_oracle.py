"""CPU tests: pin the oracle (oracle/atop_oracle.c) against
  (a) the golden vectors generated from the reference's own loop bodies and NumPy
      (tests/golden/make_golden.py), and
  (b) when oracle/_ref/libatop_ref.so is present, the reference's loop bodies live on larger
      random bit patterns.
The GPU parity tests then compare the CUDA kernels with this oracle.
"""
import zlib

import numpy as np
import pytest

from inputs import RED_N, fill_random, nan_case_input, reduction_input
from oracle import oracle as O
from util import DTYPES, FP_DTYPES, assert_bits_equal, ulp_error

REF_BIN = ("ADD", "SUB", "MUL", "DIV", "BITWISE_AND", "BITWISE_OR", "BITWISE_XOR")
REF_UN = ("ABS", "SQRT", "ISNAN", "ISINF", "ISFINITE", "FLOOR", "CEIL", "TRUNC", "INVERT", "LOGICAL_NOT")


@pytest.mark.parametrize("dt", DTYPES)
def test_oracle_elementwise_vs_golden(golden, dt):
    a, b = golden[f"in_a_{dt}"], golden[f"in_b_{dt}"]
    checked = 0
    with np.errstate(all="ignore"):
        for op in REF_BIN:
            key = f"ref_{op}_{dt}"
            if key in golden.files:
                if dt == "bool" and op in ("ADD", "MUL"):
                    pass  # canonical 0/1 inputs: raw OR/AND == logical
                assert_bits_equal(O.binary(op, a, b), golden[key], key)
                checked += 1
        for op in O.COMPARE:
            key = f"ref_CMP_{op}_{dt}"
            if key in golden.files:
                assert_bits_equal(O.compare(op, a, b), golden[key], key)
                checked += 1
        for op in REF_UN:
            key = f"ref_{op}_{dt}"
            if key in golden.files:
                assert_bits_equal(O.unary(op, a), golden[key].view(O.unary(op, a).dtype), key)
                checked += 1
        # NumPy-defined semantics (reference defective or delegating)
        if dt != "bool":
            assert_bits_equal(O.unary("NEGATIVE", a), golden[f"np_NEGATIVE_{dt}"], "NEGATIVE")
        assert_bits_equal(O.binary("MIN", a, b), golden[f"np_MIN_{dt}"], "MIN")
        assert_bits_equal(O.binary("MAX", a, b), golden[f"np_MAX_{dt}"], "MAX")
    assert checked >= 6


@pytest.mark.parametrize("op", ["SIN", "COS", "LOG", "EXP"])
def test_oracle_trig_f32_bit_exact_vs_reference_vectors(golden, op):
    """On the valid Cephes domain the restatement is 0 ULP from the reference's vector body."""
    x = golden[f"in_{op}_float32"]
    assert_bits_equal(O.unary(op, x), golden[f"ref_{op}_float32"], op)


@pytest.mark.parametrize("op", ["SIN", "COS", "LOG", "EXP"])
def test_oracle_trig_f32_edges_follow_ieee(golden, op):
    x = golden["in_edge_float32"]
    with np.errstate(all="ignore"):
        got = O.unary(op, x)
    err = ulp_error(got, golden[f"np64_{op}_edge"])
    if op in ("SIN", "COS"):
        # inside +-8192 the Cephes reduction has ~2^-24*|x| ABSOLUTE error;
        # check <= 2 ULP only where |x| <= pi or |x| > 8192 (libm path)
        m = (np.abs(x) <= 3.2) | ~(np.abs(x) <= 8192)
        assert np.all(err[m] <= 2.0), (x[m][err[m] > 2], err[m][err[m] > 2])
        assert np.all(np.abs(got[~m].astype(np.float64) - golden[f"np64_{op}_edge"][~m]) < 2e-3)
    else:
        assert np.all(err <= 2.0), (x[err > 2], got[err > 2], err[err > 2])


@pytest.mark.parametrize("dt", DTYPES)
def test_oracle_reductions_vs_golden(golden, dt):
    a = reduction_input(dt)
    assert zlib.crc32(a.tobytes()) == int(golden[f"red_crc_{dt}"]), "input generator drifted"
    with np.errstate(all="ignore"):
        s = O.reduce("ADD", a, start=np.zeros(1, a.dtype)[0])
        if a.dtype.kind == "f":
            ref = float(np.sum(a.astype(np.float64)))
            tol = (2.0 ** -23 if dt == "float32" else 20 * 2.0 ** -53) * np.abs(a.astype(np.float64)).sum()
            assert abs(float(s) - ref) <= tol
            if dt == "float64":  # the reference's own (non-defective) float64 sum: same association
                assert abs(float(s) - float(golden[f"refred_SUM_{dt}"])) <= tol
        else:
            assert s == golden[f"red_SUM_{dt}"], (s, golden[f"red_SUM_{dt}"])
            if f"refred_SUM_{dt}" in golden.files:
                assert s == golden[f"refred_SUM_{dt}"]
        first = a[0]
        assert O.reduce("MIN", a[1:], start=first) == golden[f"red_MIN_{dt}"]
        assert O.reduce("MAX", a[1:], start=first) == golden[f"red_MAX_{dt}"]
        assert O.argminmax(0, a) == int(golden[f"red_ARGMAX_{dt}"])
        assert O.argminmax(1, a) == int(golden[f"red_ARGMIN_{dt}"])


@pytest.mark.parametrize("dt", FP_DTYPES)
@pytest.mark.parametrize("tag", ["none", "first", "last", "mid"])
def test_oracle_nan_reductions_vs_golden(golden, dt, tag):
    a = nan_case_input(dt, tag)
    assert zlib.crc32(a.tobytes()) == int(golden[f"nan_crc_{tag}_{dt}"])
    for op in ("MIN", "MAX"):
        got = O.reduce(op, a[1:], start=a[0])
        want = golden[f"nan_{op}_{tag}_{dt}"]
        assert (np.isnan(got) and np.isnan(want)) or got == want, (op, got, want)
    assert O.argminmax(0, a) == int(golden[f"nan_ARGMAX_{tag}_{dt}"])
    assert O.argminmax(1, a) == int(golden[f"nan_ARGMIN_{tag}_{dt}"])


# ---- live against the reference's loop bodies (this container only) ------------------------------
needs_ref = pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref/libatop_ref.so not built")


@needs_ref
@pytest.mark.parametrize("dt", DTYPES)
def test_oracle_vs_live_reference_elementwise(rng, dt):
    n = (1 << 16) + 13  # forces the reference's scalar tail as well
    a, b = fill_random(n, dt, rng), fill_random(n, dt, rng)
    scalar = b[5:6]
    with np.errstate(all="ignore"):
        for op in REF_BIN:
            for bb in (b, np.broadcast_to(scalar, (1,))[0]):
                r = O.ref_binary(op, a, bb)
                if r is not None:
                    assert_bits_equal(O.binary(op, a, bb), r, f"{op} {dt}")
        for op in O.COMPARE:
            r = O.ref_binary(op, a, b, compare_op=True)
            if r is not None:
                assert_bits_equal(O.compare(op, a, b), r, f"cmp {op} {dt}")
        for op in REF_UN:
            if dt == "bool" and op == "ABS":
                continue
            r = O.ref_unary(op, a)
            if r is not None:
                mine = O.unary(op, a)
                assert_bits_equal(mine, r.view(mine.dtype), f"{op} {dt}")


@needs_ref
@pytest.mark.parametrize("op,lo,hi", [("SIN", -8192, 8192), ("COS", -8192, 8192), ("SIN", -3.2, 3.2),
                                      ("EXP", -87.3, 88.37), ("LOG", None, None)])
def test_oracle_trig_vs_live_reference(rng, op, lo, hi):
    n = 1 << 18  # multiple of 8: no libm tail in the reference
    if op == "LOG":
        x = fill_random(n, "float32", rng)
        x = np.abs(x)
        x = x[np.isfinite(x) & (x >= np.float32(1.1754944e-38))]
        x = x[: (x.size // 8) * 8]
    else:
        x = rng.uniform(lo, hi, n).astype(np.float32)
    assert_bits_equal(O.unary(op, x), O.ref_unary(op, x), op)


def test_reduction_input_size():
    assert RED_N > 3 * 65536


def test_oracle_convert_is_numpy_astype(rng):
    """dtype conversion: the oracle's casts against NumPy's `astype` on random BIT PATTERNS
    (NaN, inf, denormals, values far outside every integer range) for all 121 (from, to) pairs -- bit-exact, NaN to
    NaN.  One documented exception: float -> uint32 of values outside (-2^31, 2^32) (and NaN), where NumPy's own
    answer depends on whether the element falls in the vectorised body or the scalar tail of its loop (both are
    undefined behaviour in C); there the test keeps to the representable range."""
    import warnings
    from oracle import oracle as O
    for di in DTYPES:
        a = fill_random(50021, di, rng)
        if np.dtype(di).kind == "f":
            a[::3] = rng.normal(0, 1e5, a[::3].size).astype(di)
            edge = np.array([2**31, -2**31, 2**31 - 1, 2**63, -2**63, 2**64, 255.9, -0.9, 65536.5, -32769.0, 3e9,
                             4294967295.0, 4294967296.0, -1.0, 0.0, -0.0, np.inf, -np.inf, np.nan], dtype=di)
            a[1::7] = edge[np.arange(a[1::7].size) % edge.size]
        for do in DTYPES:
            x = a
            if np.dtype(di).kind == "f" and do == "uint32":
                x = a[(a > -2147483648.0) & (a < 4294967296.0)]
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                want = x.astype(do)
            assert_bits_equal(O.astype(x, do), want, f"{di} -> {do}")
