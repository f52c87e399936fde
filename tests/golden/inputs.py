"""Seeded input generators shared by make_golden.py and the tests (the fixtures store only the
expected results plus a CRC of each regenerated input)."""
import numpy as np

RED_N = 3 * 65536 + 77  # three full logical reduction blocks + a ragged one


def fill_random(n, dtype, rng):
    """Uniform random BIT PATTERNS, so floats include NaN/Inf/denormals
    (reference fixture: tests/test_ufuncs.py:36-50)."""
    dt = np.dtype(dtype)
    raw = rng.integers(0, 2**64, size=(n * dt.itemsize + 7) // 8, dtype=np.uint64)
    a = raw.view(np.uint8)[: n * dt.itemsize].view(dt).copy()
    if dt == np.bool_:
        a = (a.view(np.uint8) & 1).view(np.bool_)
    return a


def reduction_input(dtype, n=RED_N, seed=4321):
    rng = np.random.default_rng(seed)
    if np.dtype(dtype).kind == "f":
        return rng.normal(0, 100, n).astype(dtype)
    return fill_random(n, dtype, rng)


def nan_case_input(dtype, tag, n=RED_N, seed=987):
    a = np.random.default_rng(seed).normal(0, 100, n).astype(dtype)
    pos = {"none": [], "first": [0], "last": [n - 1], "mid": [70000, 150000]}[tag]
    a[pos] = np.nan
    return a
