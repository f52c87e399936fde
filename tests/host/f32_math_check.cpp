// CPU check of pnumpy_b200/csrc/atop/f32_math.cuh (the float32 asinh / acosh / atanh fast bodies of the CUDA
// kernels; the header compiles as plain C++).  Test infrastructure: built and run by tests/test_f32_math_host.py
// with  g++ -O2 -mfma -ffp-contract=off.  The hardware reciprocal / reciprocal-square-root estimates are replaced
// by correctly rounded values moved by `skew` ULP; the program sweeps skew = -2 .. 2 and reports the worst case.
// Prints one line per function: name, samples, max ULP error against double libm, worst input.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../pnumpy_b200/csrc/atop/f32_math.cuh"

static uint64_t s_rng = 0x9E3779B97F4A7C15ull;
static uint64_t rnd() { s_rng ^= s_rng << 13; s_rng ^= s_rng >> 7; s_rng ^= s_rng << 17; return s_rng; }
static float uni(double lo, double hi) { return (float)(lo + (hi - lo) * ((rnd() >> 11) * (1.0 / 9007199254740992.0))); }
static float logu(double lo, double hi) { return (float)exp(log(lo) + (log(hi) - log(lo)) * ((rnd() >> 11) * (1.0 / 9007199254740992.0))); }

static double ulp_err(float got, double want) {
    if (isnan(got) || isinf(got)) return 1e9;
    float w = (float)want;
    int e;
    frexpf(w, &e);
    double ulp = ldexp(1.0, e - 24);
    if (fabsf(w) < 1.17549435e-38f) ulp = 1.4012984643e-45;
    return fabs((double)got - want) / ulp;
}
struct Stat { double worst = 0; float at = 0; long n = 0; };
static void upd(Stat& s, float x, float got, double want) {
    double e = ulp_err(got, want);
    s.n++;
    if (e > s.worst) { s.worst = e; s.at = x; }
}

int main(int argc, char** argv) {
    long N = argc > 1 ? atol(argv[1]) : 1000000;
    Stat sa, sc, st;
    int bad = 0;
    for (int skew = -2; skew <= 2; ++skew) {
        f32_estimate_skew = skew;
        for (long i = 0; i < N; ++i) {
            float xs[6] = {uni(-80, 80), uni(-2, 2), logu(1e-30, 1e18) * ((rnd() & 1) ? 1.f : -1.f), uni(-1e-3, 1e-3),
                           logu(0.25, 4), uni(-1e6, 1e6)};
            for (float x : xs) if (f32_asinh_ok(x)) upd(sa, x, f32_asinh_fast(x), asinh((double)x));
            float ys[5] = {uni(1, 80), uni(1, 1.001), 1.0f + logu(1e-7, 1.0), logu(1, 1e18), uni(1, 3)};
            for (float x : ys) if (f32_acosh_ok(x)) upd(sc, x, f32_acosh_fast(x), acosh((double)x));
            float zs[5] = {uni(-0.9, 0.9), uni(-1, 1), logu(1e-30, 1.0), -logu(1e-30, 1.0), 1.0f - logu(6e-8, 0.5)};
            for (float x : zs) if (f32_atanh_ok(x)) upd(st, x, f32_atanh_fast(x), atanh((double)x));
        }
        // every float next to 1 and next to the ends of the domains
        for (int k = 0; k < 4096; ++k) {
            float x = f32_from(0x3f800000 + k);
            upd(sc, x, f32_acosh_fast(x), acosh((double)x));
            x = f32_from(0x3f7fffff - k);
            upd(st, x, f32_atanh_fast(x), atanh((double)x));
            upd(st, -x, f32_atanh_fast(-x), atanh(-(double)x));
            x = f32_from(0x5d7fffff - k);
            upd(sa, x, f32_asinh_fast(x), asinh((double)x));
            upd(sc, x, f32_acosh_fast(x), acosh((double)x));
        }
        // exact values
        if (f32_bits(f32_asinh_fast(0.0f)) != 0 || f32_bits(f32_asinh_fast(-0.0f)) != (int)0x80000000) { printf("asinh(+-0) wrong\n"); bad = 1; }
        if (f32_bits(f32_atanh_fast(0.0f)) != 0 || f32_bits(f32_atanh_fast(-0.0f)) != (int)0x80000000) { printf("atanh(+-0) wrong\n"); bad = 1; }
        if (f32_bits(f32_acosh_fast(1.0f)) != 0) { printf("acosh(1) wrong\n"); bad = 1; }
    }
    // domain predicates
    if (f32_asinh_ok(NAN) || f32_asinh_ok(INFINITY) || f32_asinh_ok(-INFINITY) || f32_asinh_ok(2e18f) || !f32_asinh_ok(1e18f) || !f32_asinh_ok(-0.0f)) { printf("asinh_ok wrong\n"); bad = 1; }
    if (f32_acosh_ok(NAN) || f32_acosh_ok(INFINITY) || f32_acosh_ok(0.99999994f) || f32_acosh_ok(-2.0f) || f32_acosh_ok(0.0f) || !f32_acosh_ok(1.0f) || !f32_acosh_ok(1e18f)) { printf("acosh_ok wrong\n"); bad = 1; }
    if (f32_atanh_ok(NAN) || f32_atanh_ok(1.0f) || f32_atanh_ok(-1.0f) || f32_atanh_ok(INFINITY) || !f32_atanh_ok(0.99999994f) || !f32_atanh_ok(-0.0f)) { printf("atanh_ok wrong\n"); bad = 1; }
    printf("asinh %ld %.3f %.9g\n", sa.n, sa.worst, sa.at);
    printf("acosh %ld %.3f %.9g\n", sc.n, sc.worst, sc.at);
    printf("atanh %ld %.3f %.9g\n", st.n, st.worst, st.at);
    return bad;
}
