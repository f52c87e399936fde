"""CPU check of the float64 sin/cos/exp/log fast bodies the CUDA kernels use
(pnumpy_b200/csrc/atop/f64_math.cuh compiles as plain C++): max ULP error against long double libm
on ~20M samples, exact values at 0 / -0 / 1, and the fast-domain predicates."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host", "f64_math_check.cpp")

# bounds asserted (measured: sin 1.47, cos 1.46, exp 0.93, log 0.75, log2 0.81, log10 0.91, log1p 0.73, expm1 2.0, tanh 2.5, sinh 2.4, tan 2.9);
# the contract is <= 2 ULP for sin/cos/exp/log and <= 4 ULP for the others
LIMIT = {"sin": 1.6, "cos": 1.6, "exp": 1.0, "exp2": 1.1, "log": 0.85, "log2": 0.95, "log10": 1.05, "log1p": 0.85, "expm1": 2.3,
         "tanh": 2.8, "sinh": 2.7, "tan": 3.2, "atanh": 2.1, "asinh": 1.9, "acosh": 2.7, "sqrt": 0.6,
         "atan": 2.5, "asin": 2.4, "acos": 1.4}


def test_f64_fast_bodies_against_long_double(tmp_path):
    exe = str(tmp_path / "f64chk")
    subprocess.check_call(["g++", "-O2", "-mfma", "-ffp-contract=off", "-std=c++17", "-o", exe, SRC, "-lm"])
    r = subprocess.run([exe, "1000000"], capture_output=True, text=True, timeout=300)
    sys.stdout.write(r.stdout)
    assert r.returncode == 0, r.stdout + r.stderr
    seen = {}
    for line in r.stdout.splitlines():
        f = line.split()
        if len(f) == 4 and f[0] in LIMIT:
            seen[f[0]] = (int(f[1]), float(f[2]))
    assert set(seen) == set(LIMIT)
    for name, (n, worst) in seen.items():
        assert n >= 1_000_000, (name, n)
        assert worst <= LIMIT[name], (name, worst)


def test_coefficient_file_is_what_the_generator_writes():
    """f64_coeffs.inc is generated (tools/gen_f64_coeffs.py); regenerating must reproduce it."""
    pytest.importorskip("mpmath")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_f64_coeffs.py")], capture_output=True,
                         text=True, timeout=600, check=True).stdout
    assert out == open(os.path.join(ROOT, "pnumpy_b200", "csrc", "atop", "f64_coeffs.inc")).read()
