// CPU check of pnumpy_b200/csrc/atop/f64_math.cuh (the float64 sin/cos/exp/log fast paths of the CUDA
// kernels; the header compiles as plain C++).  Test infrastructure: built and run by
// tests/test_f64_math_host.py with  g++ -O2 -mfma -ffp-contract=off.
// Prints one line per function: name, samples, max ULP error against long double libm, worst input.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../pnumpy_b200/csrc/atop/f64_math.cuh"

static uint64_t s_rng = 0x9E3779B97F4A7C15ull;
static uint64_t rnd() { s_rng ^= s_rng << 13; s_rng ^= s_rng >> 7; s_rng ^= s_rng << 17; return s_rng; }
static double uni(double lo, double hi) { return lo + (hi - lo) * ((rnd() >> 11) * (1.0 / 9007199254740992.0)); }
static double from_bits(uint64_t b) { double x; memcpy(&x, &b, 8); return x; }

static double ulp_err(double got, long double want) {
    if (isnan(got) || isinf(got)) return 1e9;
    double w = (double)want;
    int e;
    frexp(w, &e);
    long double ulp = ldexpl(1.0L, e - 53);
    if (fabs(w) < 2.2250738585072014e-308) ulp = 4.9406564584124654e-324L;
    return (double)(fabsl((long double)got - want) / ulp);
}

static double hlog(double x) { const int i = f64_log_index(x); return f64_log_fast(x, kLogTab[2 * i], kLogTab[2 * i + 1]); }

static double hlog2(double x) { const int i = f64_log_index(x); return f64_log2_fast(x, kLogTab[2 * i], kLogTab[2 * i + 1]); }
static double hlog10(double x) { const int i = f64_log_index(x); return f64_log10_fast(x, kLogTab[2 * i], kLogTab[2 * i + 1]); }
static double hlog1p(double x) { const int i = f64_log1p_index(x); return f64_log1p_fast(x, kLogTab[2 * i], kLogTab[2 * i + 1]); }

struct Stat { double worst = 0, at = 0; long n = 0; };
static void upd(Stat& s, double x, double got, long double want) {
    double e = ulp_err(got, want);
    s.n++;
    if (e > s.worst) { s.worst = e; s.at = x; }
}

int main(int argc, char** argv) {
    long N = argc > 1 ? atol(argv[1]) : 1000000;
    Stat ssin, scos, sexp, slog;
    // sin / cos: uniform in several ranges + points next to multiples of pi/2
    const double ranges[][2] = {{-1, 1}, {-20, 20}, {-8192, 8192}, {-105000, 105000}, {-1e-3, 1e-3}};
    for (auto& rg : ranges)
        for (long i = 0; i < N; ++i) {
            double x = uni(rg[0], rg[1]);
            if (!f64_sincos_ok(x)) continue;
            upd(ssin, x, f64_sincos_fast<0>(x), sinl((long double)x));
            upd(scos, x, f64_sincos_fast<1>(x), cosl((long double)x));
        }
    for (long k = -60000; k <= 60000; ++k)
        for (int d = -3; d <= 3; ++d) {
            double x = (double)(k * 1.57079632679489661923L);
            uint64_t b; memcpy(&b, &x, 8); b += (int64_t)d; x = from_bits(b);
            if (!f64_sincos_ok(x)) continue;
            upd(ssin, x, f64_sincos_fast<0>(x), sinl((long double)x));
            upd(scos, x, f64_sincos_fast<1>(x), cosl((long double)x));
        }
    // exact values that must hold
    int bad = 0;
    if (f64_sincos_fast<1>(0.0) != 1.0) { printf("cos(0) != 1\n"); bad = 1; }
    if (f64_sincos_fast<1>(1e-9) != 1.0) { printf("cos(1e-9) != 1\n"); bad = 1; }
    { double s = f64_sincos_fast<0>(-0.0); if (s != 0.0 || !signbit(s)) { printf("sin(-0) != -0\n"); bad = 1; } }
    if (f64_sincos_fast<0>(1e-300) != 1e-300) { printf("sin(1e-300) != 1e-300\n"); bad = 1; }
    if (f64_sincos_fast<0>(5e-324) != 5e-324) { printf("sin(denormal min) wrong\n"); bad = 1; }
    // exp
    const double eranges[][2] = {{-1, 1}, {-80, 80}, {-707, 707}, {-1e-5, 1e-5}};
    for (auto& rg : eranges)
        for (long i = 0; i < N; ++i) {
            double x = uni(rg[0], rg[1]);
            if (!f64_exp_ok(x)) continue;
            upd(sexp, x, f64_exp_fast(x), expl((long double)x));
        }
    if (f64_exp_fast(0.0) != 1.0 || f64_exp_fast(-0.0) != 1.0) { printf("exp(0) != 1\n"); bad = 1; }
    // exp2: the same ranges scaled, exact at integers, domain
    Stat sexp2;
    const double e2ranges[][2] = {{-1, 1}, {-100, 100}, {-1019.9, 1019.9}, {-1e-5, 1e-5}, {-0.5, 0.5}};
    for (auto& rg : e2ranges)
        for (long i = 0; i < N; ++i) {
            double x = uni(rg[0], rg[1]);
            if (!f64_exp2_ok(x)) continue;
            upd(sexp2, x, f64_exp2_fast(x), exp2l((long double)x));
        }
    for (int k = -1019; k <= 1019; ++k)
        if (f64_exp2_fast((double)k) != ldexp(1.0, k)) { printf("exp2(%d) not exact\n", k); bad = 1; break; }
    if (f64_exp2_fast(-0.0) != 1.0 || f64_exp2_ok(1020.0) || f64_exp2_ok(-1020.0) || f64_exp2_ok(NAN) || f64_exp2_ok(INFINITY) || !f64_exp2_ok(1019.99) ||
        !f64_exp2_ok(-1019.99) || !f64_exp2_ok(5e-324)) { printf("exp2 domain / exp2(-0) wrong\n"); bad = 1; }
    // log: uniform, near 1, random bit patterns of positive normals
    const double lranges[][2] = {{0.5, 2}, {1e-3, 1e6}, {0.99, 1.01}, {1 - 1e-9, 1 + 1e-9}, {0.6, 0.8}, {1.3, 1.5}};
    for (auto& rg : lranges)
        for (long i = 0; i < N; ++i) {
            double x = uni(rg[0], rg[1]);
            upd(slog, x, hlog(x), logl((long double)x));
        }
    for (long i = 0; i < N; ++i) {
        double x = from_bits((rnd() & 0x7fffffffffffffffull));
        if (!f64_log_ok(x)) continue;
        upd(slog, x, hlog(x), logl((long double)x));
    }
    if (hlog(1.0) != 0.0) { printf("log(1) != 0\n"); bad = 1; }
    if (f64_log_ok(0.0) || f64_log_ok(-1.0) || f64_log_ok(INFINITY) || f64_log_ok(NAN) || f64_log_ok(1e-310) || !f64_log_ok(2.3e-308) ||
        !f64_log_ok(1.7e308)) { printf("f64_log_ok domain wrong\n"); bad = 1; }
    if (f64_sincos_ok(INFINITY) || f64_sincos_ok(NAN) || f64_sincos_ok(2e5) || !f64_sincos_ok(-105000.0)) { printf("f64_sincos_ok domain wrong\n"); bad = 1; }
    if (f64_exp_ok(INFINITY) || f64_exp_ok(NAN) || f64_exp_ok(708.0) || f64_exp_ok(-710.0) || !f64_exp_ok(-706.9)) { printf("f64_exp_ok domain wrong\n"); bad = 1; }
    // log2 / log10: same inputs as log; exact at powers of 2 / 10
    Stat slog2, slog10, slog1p, sexpm1;
    for (auto& rg : lranges)
        for (long i = 0; i < N; ++i) {
            double x = uni(rg[0], rg[1]);
            upd(slog2, x, hlog2(x), log2l((long double)x));
            upd(slog10, x, hlog10(x), log10l((long double)x));
        }
    for (long i = 0; i < N; ++i) {
        double x = from_bits((rnd() & 0x7fffffffffffffffull));
        if (!f64_log_ok(x)) continue;
        upd(slog2, x, hlog2(x), log2l((long double)x));
        upd(slog10, x, hlog10(x), log10l((long double)x));
    }
    for (int k = -1022; k <= 1023; ++k) if (hlog2(ldexp(1.0, k)) != (double)k) { printf("log2(2^%d) inexact\n", k); bad = 1; break; }
    { double p10 = 1.0; for (int k = 0; k <= 22; ++k) { if (hlog10(p10) != (double)k) { printf("log10(1e%d) = %.17g\n", k, hlog10(p10)); bad = 1; } p10 *= 10.0; } }
    // log1p
    const double pranges[][2] = {{-0.999, 1}, {-1e-3, 1e-3}, {-1e-12, 1e-12}, {0, 1e6}, {-0.5, 0.5}, {1e6, 1e17}, {-1e-300, 1e-300}};
    for (auto& rg : pranges)
        for (long i = 0; i < N; ++i) {
            double x = uni(rg[0], rg[1]);
            if (!f64_log1p_ok(x)) continue;
            upd(slog1p, x, hlog1p(x), log1pl((long double)x));
        }
    { double z = hlog1p(-0.0); if (z != 0.0 || !signbit(z) || hlog1p(0.0) != 0.0 || signbit(hlog1p(0.0))) { printf("log1p(+-0) wrong\n"); bad = 1; } }
    if (f64_log1p_ok(-1.0) || f64_log1p_ok(-2.0) || f64_log1p_ok(INFINITY) || f64_log1p_ok(NAN) || f64_log1p_ok(1e19) || !f64_log1p_ok(-0.99) ||
        !f64_log1p_ok(1e17) || !f64_log1p_ok(5e-324) || !f64_log1p_ok(-5e-324)) { printf("f64_log1p_ok domain wrong\n"); bad = 1; }
    // expm1
    const double mranges[][2] = {{-1, 1}, {-40, 40}, {-707, 707}, {-1e-5, 1e-5}, {-0.4, 0.4}, {-1e-300, 1e-300}};
    for (auto& rg : mranges)
        for (long i = 0; i < N; ++i) {
            double x = uni(rg[0], rg[1]);
            if (!f64_exp_ok(x)) continue;
            upd(sexpm1, x, f64_expm1_fast(x), expm1l((long double)x));
        }
    { double z = f64_expm1_fast(-0.0); if (z != 0.0 || !signbit(z)) { printf("expm1(-0) wrong\n"); bad = 1; } }
    // tanh / sinh / tan
    Stat stanh, ssinh, stan;
    const double hranges[][2] = {{-1, 1}, {-20, 20}, {-350, 350}, {-1e-5, 1e-5}, {-0.5, 0.5}, {-1e-300, 1e-300}};
    for (auto& rg : hranges)
        for (long i = 0; i < N; ++i) {
            double x = uni(rg[0], rg[1]);
            if (f64_tanh_ok(x)) upd(stanh, x, f64_tanh_fast(x), tanhl((long double)x));
            if (f64_exp_ok(x)) upd(ssinh, x, f64_sinh_fast(x), sinhl((long double)x));
        }
    for (auto& rg : ranges)
        for (long i = 0; i < N; ++i) {
            double x = uni(rg[0], rg[1]);
            if (!f64_sincos_ok(x)) continue;
            upd(stan, x, f64_tan_fast(x), tanl((long double)x));
        }
    for (long k = -60000; k <= 60000; ++k)
        for (int d = -3; d <= 3; ++d) {
            double x = (double)(k * 1.57079632679489661923L);
            uint64_t b; memcpy(&b, &x, 8); b += (int64_t)d; x = from_bits(b);
            if (!f64_sincos_ok(x)) continue;
            upd(stan, x, f64_tan_fast(x), tanl((long double)x));
        }
    if (f64_tanh_fast(40.0) != 1.0 || f64_tanh_fast(-40.0) != -1.0) { printf("tanh(+-40) != +-1\n"); bad = 1; }
    { double z = f64_tanh_fast(-0.0); if (z != 0.0 || !signbit(z)) { printf("tanh(-0) wrong\n"); bad = 1; } }
    { double z = f64_sinh_fast(-0.0); if (z != 0.0 || !signbit(z)) { printf("sinh(-0) wrong\n"); bad = 1; } }
    { double z = f64_tan_fast(-0.0); if (z != 0.0 || !signbit(z)) { printf("tan(-0) wrong\n"); bad = 1; } }
    if (f64_tanh_fast(1e-300) != 1e-300 || f64_sinh_fast(1e-300) != 1e-300 || f64_tan_fast(1e-300) != 1e-300) { printf("tiny argument wrong\n"); bad = 1; }
    // atanh / asinh / acosh
    Stat satanh, sasinh, sacosh;
    F64FlatLogTable flat;
    const double aranges[][2] = {{-0.999, 0.999}, {-0.5, 0.5}, {-1e-4, 1e-4}, {-1e-300, 1e-300}};
    for (auto& rg : aranges)
        for (long i = 0; i < N; ++i) {
            double x = uni(rg[0], rg[1]);
            if (f64_atanh_ok(x)) upd(satanh, x, f64_atanh_fast(x, flat), atanhl((long double)x));
        }
    const double sranges[][2] = {{-1, 1}, {-100, 100}, {-1e8, 1e8}, {-1e-4, 1e-4}, {-1e-300, 1e-300}, {-3, 3}};
    for (auto& rg : sranges)
        for (long i = 0; i < N; ++i) {
            double x = uni(rg[0], rg[1]);
            if (f64_asinh_ok(x)) upd(sasinh, x, f64_asinh_fast(x, flat), asinhl((long double)x));
        }
    const double cranges[][2] = {{1, 2}, {1, 100}, {1, 1e8}, {1, 1.0001}, {1, 1.0000000001}};
    for (auto& rg : cranges)
        for (long i = 0; i < N; ++i) {
            double x = uni(rg[0], rg[1]);
            if (f64_acosh_ok(x)) upd(sacosh, x, f64_acosh_fast(x, flat), acoshl((long double)x));
        }
    if (f64_acosh_fast(1.0, flat) != 0.0) { printf("acosh(1) != 0\n"); bad = 1; }
    if (f64_acosh_ok(0.999) || f64_acosh_ok(-2.0) || f64_acosh_ok(NAN) || f64_acosh_ok(1e9) || !f64_acosh_ok(1.0) || !f64_acosh_ok(1e8)) { printf("f64_acosh_ok domain wrong\n"); bad = 1; }
    { double z = f64_atanh_fast(-0.0, flat); if (z != 0.0 || !signbit(z)) { printf("atanh(-0) wrong\n"); bad = 1; } }
    { double z = f64_asinh_fast(-0.0, flat); if (z != 0.0 || !signbit(z)) { printf("asinh(-0) wrong\n"); bad = 1; } }
    { Stat ssq; for (long i = 0; i < N; ++i) { double v = uni(1e-3, 1e6); upd(ssq, v, f64_sqrt(v), sqrtl((long double)v)); } printf("sqrt %ld %.4f %.17g\n", ssq.n, ssq.worst, ssq.at); }
    // atan / asin / acos
    Stat satan, sasin, sacos;
    const double tranges[][2] = {{-1, 1}, {-0.45, 0.45}, {-3, 3}, {-100, 100}, {-1e6, 1e6}, {-1e-5, 1e-5}, {-1e-300, 1e-300}};
    for (auto& rg : tranges)
        for (long i = 0; i < N; ++i) {
            double x = uni(rg[0], rg[1]);
            upd(satan, x, f64_atan_fast(x), atanl((long double)x));
        }
    const double iranges[][2] = {{-0.999999, 0.999999}, {-0.5, 0.5}, {0.4, 0.6}, {0.99, 0.9999999}, {-1e-5, 1e-5}, {-1e-300, 1e-300}};
    for (auto& rg : iranges)
        for (long i = 0; i < N; ++i) {
            double x = uni(rg[0], rg[1]);
            if (!f64_asin_ok(x)) continue;
            upd(sasin, x, f64_asin_fast(x), asinl((long double)x));
            upd(sacos, x, f64_acos_fast(x), acosl((long double)x));
        }
    { double z = f64_atan_fast(-0.0); if (z != 0.0 || !signbit(z)) { printf("atan(-0) wrong\n"); bad = 1; } }
    { double z = f64_asin_fast(-0.0); if (z != 0.0 || !signbit(z)) { printf("asin(-0) wrong\n"); bad = 1; } }
    if (f64_atan_fast(1e300) != 1.5707963267948966 || f64_atan_fast(-1e300) != -1.5707963267948966) { printf("atan(huge) wrong\n"); bad = 1; }
    if (f64_atan_fast(1.0) != 0.7853981633974483) { printf("atan(1) = %.17g\n", f64_atan_fast(1.0)); bad = 1; }
    if (f64_acos_fast(0.0) != 1.5707963267948966) { printf("acos(0) = %.17g\n", f64_acos_fast(0.0)); bad = 1; }
    if (f64_asin_ok(1.0) || f64_asin_ok(-1.0) || f64_asin_ok(NAN) || f64_asin_ok(2.0) || !f64_asin_ok(0.9999999999)) { printf("f64_asin_ok domain wrong\n"); bad = 1; }
    printf("atan %ld %.4f %.17g\n", satan.n, satan.worst, satan.at);
    printf("asin %ld %.4f %.17g\n", sasin.n, sasin.worst, sasin.at);
    printf("acos %ld %.4f %.17g\n", sacos.n, sacos.worst, sacos.at);
    printf("atanh %ld %.4f %.17g\n", satanh.n, satanh.worst, satanh.at);
    printf("asinh %ld %.4f %.17g\n", sasinh.n, sasinh.worst, sasinh.at);
    printf("acosh %ld %.4f %.17g\n", sacosh.n, sacosh.worst, sacosh.at);
    printf("tanh %ld %.4f %.17g\n", stanh.n, stanh.worst, stanh.at);
    printf("sinh %ld %.4f %.17g\n", ssinh.n, ssinh.worst, ssinh.at);
    printf("tan %ld %.4f %.17g\n", stan.n, stan.worst, stan.at);
    printf("log2 %ld %.4f %.17g\n", slog2.n, slog2.worst, slog2.at);
    printf("log10 %ld %.4f %.17g\n", slog10.n, slog10.worst, slog10.at);
    printf("log1p %ld %.4f %.17g\n", slog1p.n, slog1p.worst, slog1p.at);
    printf("expm1 %ld %.4f %.17g\n", sexpm1.n, sexpm1.worst, sexpm1.at);
    printf("sin %ld %.4f %.17g\n", ssin.n, ssin.worst, ssin.at);
    printf("cos %ld %.4f %.17g\n", scos.n, scos.worst, scos.at);
    printf("exp %ld %.4f %.17g\n", sexp.n, sexp.worst, sexp.at);
    printf("exp2 %ld %.4f %.17g\n", sexp2.n, sexp2.worst, sexp2.at);
    printf("log %ld %.4f %.17g\n", slog.n, slog.worst, slog.at);
    return bad;
}
