// float32 asinh / acosh / atanh for the streaming kernels (ops_trig_more.cu).
//
// Why not CUDA's asinhf()/acoshf()/atanhf(): inside the streaming kernel they cost 50-70 issue slots per
// element (a full logf with its own special-case branches, an IEEE division and an IEEE square root with
// their slow-path tests) and the kernels ran at 4.2 / 5.6 / 6.3 TB/s on B200 (profiles/r01_transcendentals.md),
// i.e. issue-bound, not HBM-bound: at 6.4 TB/s (80 % of the nominal 8 TB/s) an SM retires one float32 element
// every 0.35 clocks, which leaves ~45 issue slots per element.
//
// These bodies cost 28-33 FP32/integer instructions + 2-3 MUFU (separate pipe):
//   * one shared log1p core: u = fl(1 + r), c = r - (u - 1) (the rounding error of u, exact), u = 2^e z with
//     z in [sqrt(1/2), sqrt(2)) by integer arithmetic on the bits, log(z) = f - f^2/2 + f^3 P(f) with f = z - 1
//     (degree-7 P, Remez fit relative to the result: 2^-27.3), result = e ln2_hi + (f + (f^2 Q(f) + e ln2_lo + c/u));
//     c/u uses the hardware reciprocal estimate (c is below ulp(u)/2, so 1 ULP of 1/u is nothing);
//   * asinh(x) = sign(x) log1p(a + a^2 / (1 + sqrt(1 + a^2))), a = |x|: MUFU.RSQ for the square root (one
//     multiply, no Newton step: its 2 ULP land on a term that is at most 0.3 of the log1p argument),
//     MUFU.RCP for the division;
//   * acosh(x) = log1p(d + sqrt(d (d + 2))), d = x - 1: MUFU.RSQ + ONE Newton step (near x = 1 the result
//     IS the square root, so it has to be good to half an ULP);
//   * atanh(x) = sign(x) 1/2 log1p(2a / (1 - a)): MUFU.RCP + one Newton step on the quotient.
// No per-element branch: the kernel tests a thread's whole set of inputs once (ok()) and runs fast() on all of
// them; a thread with any input outside the fast domain (NaN, +-inf, |x| >= 2^60, acosh of x < 1, atanh of
// |x| >= 1) runs CUDA's function element by element -- a value's result never depends on its neighbours.
// Accuracy: tests/test_f32_math_host.py (this header compiles as plain C++, with the hardware estimates
// replaced by correctly rounded ones perturbed by up to +-2 ULP) and the GPU parity tests (<= 4 ULP).
// Coefficients: tools/gen_f32_coeffs.py (mpmath Remez), pasted below with the generator's own digits.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define F32_FN static __device__ __forceinline__
F32_FN float f32_fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
F32_FN float f32_add(float a, float b) { return __fadd_rn(a, b); }   // never contracted (the file is -fmad=false anyway)
F32_FN float f32_mul(float a, float b) { return __fmul_rn(a, b); }
F32_FN int f32_bits(float x) { return __float_as_int(x); }
F32_FN float f32_from(int b) { return __int_as_float(b); }
F32_FN float f32_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
F32_FN float f32_rsq(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
#else
#define F32_FN static inline
F32_FN float f32_fma(float a, float b, float c) { return fmaf(a, b, c); }
F32_FN float f32_add(float a, float b) { return a + b; }
F32_FN float f32_mul(float a, float b) { return a * b; }
F32_FN int f32_bits(float x) { int b; memcpy(&b, &x, 4); return b; }
F32_FN float f32_from(int b) { float x; memcpy(&x, &b, 4); return x; }
// CPU stand-ins for the hardware estimates: correctly rounded, then moved by f32_estimate_skew ULP.  The test sweeps
// -2 .. 2 for MUFU.RSQ (documented: 2^-22.9 relative) and -1 .. 1 for MUFU.RCP (documented: 1 ULP; exact at powers
// of two, which the tiny-argument case relies on: log1p(r) = r / 1 for r < 2^-24 -- checked on the GPU by the parity test).
static int f32_estimate_skew = 0;
F32_FN float f32_rcp(float x) {
    const float y = 1.0f / x;
    if ((f32_bits(x) & 0x007fffff) == 0) return y;
    return f32_from(f32_bits(y) + (f32_estimate_skew > 1 ? 1 : (f32_estimate_skew < -1 ? -1 : f32_estimate_skew)));
}
F32_FN float f32_rsq(float x) { return f32_from(f32_bits((float)(1.0 / sqrt((double)x))) + f32_estimate_skew); }
#endif

// log(1 + f) = f - f^2/2 + f^3 P(f) on [sqrt(1/2) - 1, sqrt(2) - 1]; kLog1pP[i] multiplies f^i
#define F32_LOG1P_P0  0.33333331707789743f
#define F32_LOG1P_P1 -0.25000821028533193f
#define F32_LOG1P_P2  0.200012268730962f
#define F32_LOG1P_P3 -0.16623357534961455f
#define F32_LOG1P_P4  0.1420175799828562f
#define F32_LOG1P_P5 -0.1316018046633975f
#define F32_LOG1P_P6  0.12761577348054695f
#define F32_LOG1P_P7 -0.0763450290965327f
#define F32_LN2_HI 0.693145751953125f          /* 0x3f317200: 15 significant bits, e * hi exact for |e| < 2^9 */
#define F32_LN2_LO 1.42860682030941723212e-6f  /* ln 2 - hi */

// log1p(r) * scale for 0 <= r < 2^100 (scale = 1 or 1/2, a power of two: folded into the constants' uses)
F32_FN float f32_log1p_pos(float r) {
    const float u = f32_add(1.0f, r);
    const float c = f32_add(r, -f32_add(u, -1.0f));        // exact for r < 2^24; beyond, c/u < 2^-24 of a result > 16
    const int iu = f32_bits(u);
    const int t = iu - 0x3f3504f3;                          // bits(u) - bits(sqrt(1/2))
    const int e = t >> 23;                                  // u = 2^e z, z in [sqrt(1/2), sqrt(2))
    const float z = f32_from(iu - (t & (int)0xff800000));
    const float f = f32_add(z, -1.0f);                      // exact (Sterbenz)
    const float ef = f32_add(f32_from(e + 0x4b400000), -12582912.0f);   // (float)e by a magic-number add: 1.5 2^23 + e
    float p = F32_LOG1P_P7;
    p = f32_fma(p, f, F32_LOG1P_P6);
    p = f32_fma(p, f, F32_LOG1P_P5);
    p = f32_fma(p, f, F32_LOG1P_P4);
    p = f32_fma(p, f, F32_LOG1P_P3);
    p = f32_fma(p, f, F32_LOG1P_P2);
    p = f32_fma(p, f, F32_LOG1P_P1);
    p = f32_fma(p, f, F32_LOG1P_P0);
    const float q = f32_fma(p, f, -0.5f);                   // -1/2 + f P(f)
    const float lo = f32_fma(ef, F32_LN2_LO, f32_mul(c, f32_rcp(u)));
    const float mid = f32_add(f, f32_fma(f32_mul(f, f), q, lo));
    return f32_fma(ef, F32_LN2_HI, mid);
}

F32_FN float f32_copysign_bits(float mag, float sgn) { return f32_from(f32_bits(mag) | (f32_bits(sgn) & (int)0x80000000)); }

// asinh: |x| < 2^60 (a^2 must not overflow); NaN and inf fail the test
F32_FN bool f32_asinh_ok(float x) { return (unsigned)(f32_bits(x) & 0x7fffffff) < 0x5d800000u; }
F32_FN float f32_asinh_fast(float x) {
    const float a = fabsf(x);
    const float a2 = f32_mul(a, a);
    const float h = f32_add(1.0f, a2);
    const float s = f32_mul(h, f32_rsq(h));                 // sqrt(1 + a^2) to ~2 ULP
    const float r = f32_fma(a2, f32_rcp(f32_add(1.0f, s)), a);
    return f32_copysign_bits(f32_log1p_pos(r), x);
}

// acosh: 1 <= x < 2^60
F32_FN bool f32_acosh_ok(float x) { return (unsigned)f32_bits(x) - 0x3f800000u < 0x5d800000u - 0x3f800000u; }
F32_FN float f32_acosh_fast(float x) {
    const float d = f32_add(x, -1.0f);                      // exact for x <= 2
    const float v = f32_fma(d, d, f32_add(d, d));           // d (d + 2) = x^2 - 1 without cancellation
    const float y = f32_rsq(fmaxf(v, 1e-37f));              // v = 0 (x = 1): 0 * finite = 0, not 0 * inf
    const float s0 = f32_mul(v, y);
    const float s = f32_fma(f32_fma(-s0, s0, v), f32_mul(0.5f, y), s0);   // one Newton step with the exact residual
    return f32_log1p_pos(f32_add(d, s));
}

// atanh: |x| < 1
F32_FN bool f32_atanh_ok(float x) { return (unsigned)(f32_bits(x) & 0x7fffffff) < 0x3f800000u; }
F32_FN float f32_atanh_fast(float x) {
    const float a = fabsf(x);
    const float n = f32_add(a, a);
    const float d = f32_add(1.0f, -a);                      // exact for a >= 1/2
    const float y = f32_rcp(d);
    const float q0 = f32_mul(n, y);
    const float q = f32_fma(f32_fma(-q0, d, n), y, q0);     // one correction with the exact residual
    return f32_copysign_bits(f32_mul(0.5f, f32_log1p_pos(q)), x);
}
