"""Derives the polynomial of pnumpy_b200/csrc/atop/f32_math.cuh (float32 log1p core of asinh / acosh / atanh):

    log(1 + f) = f - f^2/2 + f^3 P(f),  f in [sqrt(1/2) - 1, sqrt(2) - 1],  P of degree 7,

Remez exchange in mpmath, error weighted relative to the result.  Prints the #define lines of the header
(tests/test_f32_math_host.py checks that the header carries exactly these digits).

    python tools/gen_f32_coeffs.py
"""
import os
import sys

import mpmath as mp

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gen_f64_coeffs import remez  # noqa: E402

mp.mp.dps = 40


def lines():
    lo, hi = mp.sqrt(mp.mpf(1) / 2) - 1, mp.sqrt(2) - 1
    tiny = mp.mpf(10) ** -12

    def P(f):
        return mp.mpf(1) / 3 if abs(f) < tiny else (mp.log(1 + f) - f + f * f / 2) / f ** 3

    def w(f):
        return mp.mpf(10) ** -24 if abs(f) < tiny else abs(f ** 3 / mp.log(1 + f))

    c, err = remez(P, lo, hi, 8, weight=w, grid=2000)
    out = [f"#define F32_LOG1P_P{i} {' ' if x >= 0 else ''}{float(x)!r}f" for i, x in enumerate(c)]
    return out, float(mp.log(err, 2))


if __name__ == "__main__":
    ls, lg = lines()
    print("\n".join(ls))
    print(f"// max relative error 2^{lg:.2f}", file=sys.stderr)
