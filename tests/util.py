"""Shared helpers for the parity tests."""
import numpy as np

DTYPES = ["bool", "int8", "uint8", "int16", "uint16", "int32", "uint32", "int64", "uint64",
          "float32", "float64"]
INT_DTYPES = DTYPES[1:9]
FP_DTYPES = ["float32", "float64"]


def assert_bits_equal(got, want, what=""):
    """Bit-exact for every non-NaN value; NaN must match NaN (payload/sign unspecified:
    x86 propagates operand payloads, CUDA returns the canonical NaN -- SURVEY.md App. C.1)."""
    got, want = np.asarray(got), np.asarray(want)
    assert got.dtype == want.dtype, f"{what}: dtype {got.dtype} != {want.dtype}"
    assert got.shape == want.shape, f"{what}: shape {got.shape} != {want.shape}"
    if got.dtype.kind == "f":
        gn, wn = np.isnan(got), np.isnan(want)
        assert np.array_equal(gn, wn), f"{what}: NaN positions differ at {np.flatnonzero(gn != wn)[:5]}"
        ui = {4: np.uint32, 8: np.uint64}[got.dtype.itemsize]
        g, w = got.view(ui)[~gn], want.view(ui)[~wn]
        bad = np.flatnonzero(g != w)
        assert bad.size == 0, (f"{what}: {bad.size} values differ bitwise, first at {bad[:5]}: "
                               f"{got[~gn][bad[:5]]} vs {want[~wn][bad[:5]]}")
    else:
        bad = np.flatnonzero(got.reshape(-1).view(np.uint8) != want.reshape(-1).view(np.uint8))
        assert bad.size == 0, f"{what}: {bad.size} bytes differ, first at {bad[:5]}"


def ulp_error(got, truth64):
    """|got - truth| in units of the last place of got's dtype at truth; inf/nan must agree."""
    got = np.asarray(got)
    t = np.asarray(truth64, dtype=np.float64)
    g = got.astype(np.float64)
    special = ~np.isfinite(t) | ~np.isfinite(g)
    with np.errstate(all="ignore"):
        same_special = (np.isnan(t) & np.isnan(g)) | (t == g)
        tt = np.abs(t).astype(got.dtype)
        tiny = np.finfo(got.dtype).smallest_subnormal
        sp = np.maximum(np.spacing(tt).astype(np.float64), float(tiny))
        # a truth beyond the dtype's range rounds to inf: got == inf is then exact
        over = np.isinf(tt) & np.isinf(g) & (np.sign(g) == np.sign(t))
        err = np.abs(g - t) / sp
    err = np.where(special, np.where(same_special | over, 0.0, np.inf), err)
    return err
