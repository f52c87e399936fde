"""Generate the golden fixtures under tests/golden/ (run in the build container, where
/root/reference is mounted):

    python tests/golden/make_golden.py

Sources of truth
  * oracle/_ref/libatop_ref.so -- the reference's own AVX2 loop bodies compiled unmodified from
    /root/reference/src/atop (oracle/Makefile).  Used for every op where the reference is not
    defective: out-type / fast-pointer table of its getters, integer and float
    add/sub/mul/div, bitwise, comparisons, abs/sqrt/isnan/..., float32 sin/cos/log/exp on the
    valid Cephes domain, integer and float64 reductions.
  * NumPy (version recorded in the fixture) -- for the semantics the reference delegates to
    NumPy or gets wrong: NaN-propagating min/max, argmin/argmax, float negative, and the
    transcendental edge cases.
The fixtures are what travels to the GPU box; /root/reference does not.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import oracle as O  # noqa: E402

DTYPES = ["bool", "int8", "uint8", "int16", "uint16", "int32", "uint32", "int64", "uint64",
          "float32", "float64"]


sys.path.insert(0, HERE)
from inputs import fill_random, reduction_input, nan_case_input, RED_N  # noqa: E402


def main():
    assert O.have_ref(), "build oracle/_ref first: make -C oracle"
    rng = np.random.default_rng(1234)  # reference fixture seed: tests/conftest.py:21-30
    N = 1000  # multiple of 8: the reference's vector body covers every element

    with open(os.path.join(HERE, "getter_table.json"), "w") as f:
        json.dump(O.ref_getter_table(), f, indent=0, sort_keys=True)

    g = {}
    # ---- element-wise, from the reference's loop bodies --------------------------------------
    for dt in DTYPES:
        a, b = fill_random(N, dt, rng), fill_random(N, dt, rng)
        g[f"in_a_{dt}"], g[f"in_b_{dt}"] = a, b
        for op in ("ADD", "SUB", "MUL", "DIV", "BITWISE_AND", "BITWISE_OR", "BITWISE_XOR"):
            r = O.ref_binary(op, a, b)
            if r is not None:
                g[f"ref_{op}_{dt}"] = r
        for op in O.COMPARE:
            r = O.ref_binary(op, a, b, compare_op=True)
            if r is not None:
                g[f"ref_CMP_{op}_{dt}"] = r
        for op in ("ABS", "SQRT", "ISNAN", "ISINF", "ISFINITE", "FLOOR", "CEIL", "TRUNC",
                   "INVERT", "LOGICAL_NOT"):
            if dt == "bool" and op == "ABS":
                continue
            r = O.ref_unary(op, a)
            if r is not None:
                g[f"ref_{op}_{dt}"] = r
        # NumPy-defined semantics
        with np.errstate(all="ignore"):
            if dt != "bool":
                g[f"np_NEGATIVE_{dt}"] = np.negative(a)
            g[f"np_MIN_{dt}"] = np.minimum(a, b)
            g[f"np_MAX_{dt}"] = np.maximum(a, b)
    # ---- float32 transcendental bodies on the valid Cephes domain -------------------------------
    dom = {
        "SIN": rng.uniform(-8192, 8192, N).astype(np.float32),
        "COS": rng.uniform(-8192, 8192, N).astype(np.float32),
        "LOG": np.exp(rng.uniform(-87, 88, N)).astype(np.float32),
        "EXP": rng.uniform(-87.3, 88.37, N).astype(np.float32),
    }
    for op, x in dom.items():
        g[f"in_{op}_float32"] = x
        g[f"ref_{op}_float32"] = O.ref_unary(op, x)
    # edge cases follow NumPy (the reference is defective there)
    edge = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1e-45, 1.1754942e-38, 1.1754944e-38,
                     3.4028235e38, 1.0, -1.0, 88.3762626647949, 88.5, 88.72, 88.73, 100.0, -87.0,
                     -87.5, -88.38, -100.0, -103.0, -104.0, -110.0, 8192.0, 8193.0, 1e5, 1e7, 1e20,
                     2.0**24, 0.5, 3.1415927, 1e-30, 1e-40, -1e-40], dtype=np.float32)
    g["in_edge_float32"] = edge
    with np.errstate(all="ignore"):
        for op, fn in (("SIN", np.sin), ("COS", np.cos), ("LOG", np.log), ("EXP", np.exp)):
            g[f"np64_{op}_edge"] = fn(edge.astype(np.float64))  # correctly rounded witness
    # ---- reductions: inputs are regenerated from a seed (tests/golden/inputs.py); only the
    # expected scalars and a checksum of each input are stored ----------------------------------
    import zlib
    for dt in DTYPES:
        a = reduction_input(dt)
        g[f"red_crc_{dt}"] = np.asarray(zlib.crc32(a.tobytes()))
        with np.errstate(all="ignore"):
            if dt == "bool":
                g[f"red_SUM_{dt}"] = np.asarray(a.any())
                g[f"red_MIN_{dt}"] = np.asarray(a.all())
                g[f"red_MAX_{dt}"] = np.asarray(a.any())
            else:
                g[f"red_SUM_{dt}"] = np.asarray(np.add.reduce(a, dtype=a.dtype))
                g[f"red_MIN_{dt}"] = np.asarray(a.min())
                g[f"red_MAX_{dt}"] = np.asarray(a.max())
            g[f"red_ARGMAX_{dt}"] = np.asarray(a.argmax())
            g[f"red_ARGMIN_{dt}"] = np.asarray(a.argmin())
        if dt not in ("float32",):  # the reference's float32 sum reads out of bounds
            start = np.zeros(1, dtype=a.dtype)[0]
            r = O.ref_reduce("ADD", a, start)
            if r is not None:
                g[f"refred_SUM_{dt}"] = np.asarray(r)
    # NaN placement cases for float min/max/arg* 
    for dt in ("float32", "float64"):
        for tag in ("none", "first", "last", "mid"):
            a = nan_case_input(dt, tag)
            g[f"nan_crc_{tag}_{dt}"] = np.asarray(zlib.crc32(a.tobytes()))
            with np.errstate(all="ignore"):
                g[f"nan_MIN_{tag}_{dt}"] = np.asarray(a.min())
                g[f"nan_MAX_{tag}_{dt}"] = np.asarray(a.max())
                g[f"nan_ARGMIN_{tag}_{dt}"] = np.asarray(a.argmin())
                g[f"nan_ARGMAX_{tag}_{dt}"] = np.asarray(a.argmax())
    g["numpy_version"] = np.asarray(np.__version__)
    np.savez_compressed(os.path.join(HERE, "golden.npz"), **g)
    print("wrote", len(g), "arrays;", os.path.getsize(os.path.join(HERE, "golden.npz")) // 1024, "KiB")


if __name__ == "__main__" and "--hook-keys" not in sys.argv:
    main()


def hook_keys():
    """(ufunc name, dtype num) pairs the reference hooks at initialize() -- its atop_info() keys --
    taken from the reference's own extension built by oracle/build_ref_ext.sh.  Run in a fresh
    process (python tests/golden/make_golden.py --hook-keys): it patches NumPy's loop tables."""
    sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle", "_ref"))
    import pnumpy_ref
    pnumpy_ref.initialize()
    keys = sorted([k[0], int(k[1])] for k in pnumpy_ref.atop_info().keys())
    with open(os.path.join(HERE, "hook_keys.json"), "w") as f:
        json.dump(keys, f)
    print("wrote", len(keys), "hook keys")


if __name__ == "__main__" and "--hook-keys" in sys.argv:
    hook_keys()
