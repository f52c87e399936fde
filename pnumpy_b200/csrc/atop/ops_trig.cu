// Transcendental getters: pncu_GetTrigOpFast  (sin, cos, log, exp on float32/float64).
// replaces: GetTrigOpFast, src/atop/ops_trig.cpp:578-781, and the Cephes-derived bodies
// log256_ps / exp256_ps / sin256_ps / cos256_ps, src/atop/avx_mathfun.h:177-248, 354-409,
// 506-601, 604-673.
//
// float32: the kernels restate the Cephes operation SEQUENCE literally with round-to-nearest
// mul/add/sub intrinsics (never contracted into FMA), so on the domain where the reference's
// vector body is valid the results are bit-identical to it:
//     log: every positive normal x;  exp: x in [-87.33, 88.376];  sin/cos: |x| <= 8192
// (the range avx_mathfun.h:497-500 documents).  Outside that domain the reference is
// defective (log(0)=nan, log(inf)=88.7, exp(>88.38) clamped, denormal
// inputs/outputs flushed, garbage for huge sin arguments) and the kernels follow IEEE /
// NumPy instead: log(+-0) = -inf, log(<0) = nan, log(inf) = inf, denormal inputs rescaled,
// exp overflow -> inf, exp underflow through correctly rounded denormals, |x| > 8192 or
// non-finite sin/cos through CUDA's full-range sinf/cosf.
//
// float64: the reference's pd bodies are wrong (exponent-bias typo avx_mathfun.h:107, float
// clamp :326-327, float coefficients :472-481), so there is nothing to match; the kernels use
// CUDA's double-precision sin/cos/log/exp (<= 2 ULP, documented by NVIDIA) and are checked
// against correctly rounded values.
#include <float.h>
#include <math.h>
#include "ew.cuh"
#include "functors.cuh"

using namespace pncu;

namespace {

// rn-only arithmetic helpers
#define FM(a, b) __fmul_rn((a), (b))
#define FA(a, b) __fadd_rn((a), (b))
#define FS(a, b) __fsub_rn((a), (b))

// The kernels are instruction-issue bound, not HBM bound (about 45 issue slots per element in a
// literal transcription, profiles/r01_sweep.md), so each function has a lean FAST path that is the
// same arithmetic with the non-arithmetic glue made cheap, and a general path for the rare rest:
//   * float<->int conversions (F2I/I2F/FRND issue at quarter rate) are replaced by exact
//     "magic number" adds: for 0 <= v < 2^22, __fadd_rz(v, 2^23) = 2^23 + trunc(v) exactly, its low
//     mantissa bits ARE the integer, and (2^23 + j) - 2^23 = (float)j exactly;
//   * sin/cos evaluate ONE Horner chain with per-lane selected coefficients instead of both
//     polynomials; every retained operation has the same operands in the same order as the
//     reference's, so the rounded results are identical (the discarded polynomial never
//     influenced the result);
//   * all edge-case tests collapse into one range test that sends the lane to the general path.

// ---- log ---------------------------------------------------------------------------------
// avx_mathfun.h:177-248.  `x` positive, finite and normal on entry; eadj = 0, or -23 for a rescaled denormal.
__device__ __forceinline__ float cephes_log_core(float x, float ebias) {
    const unsigned bits = __float_as_uint(x);
    // (float)((bits >> 23) - 127) + 1 and the later "- 1 if mantissa < sqrt(1/2)" are sums of small
    // integers, exact in any order: fold them into one subtraction from the magic value 2^23 + E.
    const float emagic = __uint_as_float(0x4B000000u | (bits >> 23));
    x = __uint_as_float((bits & 0x807fffffu) | 0x3f000000u);  // mantissa in [0.5, 1)
    const bool lt = x < 0.707106781186547524f;                // cephes_SQRTHF
    const float e = FS(emagic, lt ? FA(8388735.0f, ebias) : FA(8388734.0f, ebias));
    const float tmp0 = lt ? x : 0.0f;
    x = FS(x, 1.0f);
    x = FA(x, tmp0);
    const float z = FM(x, x);
    float y = 7.0376836292E-2f;
    y = FM(y, x); y = FA(y, -1.1514610310E-1f);
    y = FM(y, x); y = FA(y, 1.1676998740E-1f);
    y = FM(y, x); y = FA(y, -1.2420140846E-1f);
    y = FM(y, x); y = FA(y, +1.4249322787E-1f);
    y = FM(y, x); y = FA(y, -1.6668057665E-1f);
    y = FM(y, x); y = FA(y, +2.0000714765E-1f);
    y = FM(y, x); y = FA(y, -2.4999993993E-1f);
    y = FM(y, x); y = FA(y, +3.3333331174E-1f);
    y = FM(y, x);
    y = FM(y, z);
    float tmp = FM(e, -2.12194440e-4f);  // cephes_log_q1
    y = FA(y, tmp);
    tmp = FM(z, 0.5f);
    y = FS(y, tmp);
    tmp = FM(e, 0.693359375f);  // cephes_log_q2
    x = FA(x, y);
    x = FA(x, tmp);
    return x;
}
__device__ __noinline__ float cephes_log_general(float x) {
    if (!(x > 0.0f)) return (x == 0.0f) ? -INFINITY : NAN;  // +-0 -> -inf; negative, NaN -> NaN
    if (x == INFINITY) return INFINITY;
    // denormal: rescale exactly by 2^23 (the reference clamps to FLT_MIN instead)
    return cephes_log_core(FM(x, 8388608.0f), 23.0f);
}
__device__ __forceinline__ float cephes_logf(float x) {
    // positive normal finite <=> bits in [0x00800000, 0x7f800000)
    if (__float_as_uint(x) - 0x00800000u >= 0x7f000000u) return cephes_log_general(x);
    return cephes_log_core(x, 0.0f);
}

// ---- exp ---------------------------------------------------------------------------------
// avx_mathfun.h:354-409
__device__ __forceinline__ float cephes_exp_poly(float x, float fx) {
    float tmp = FM(fx, 0.693359375f);             // cephes_exp_C1
    float z = FM(fx, -2.12194440e-4f);            // cephes_exp_C2
    x = FS(x, tmp);
    x = FS(x, z);
    z = FM(x, x);
    float y = 1.9875691500E-4f;
    y = FM(y, x); y = FA(y, 1.3981999507E-3f);
    y = FM(y, x); y = FA(y, 8.3334519073E-3f);
    y = FM(y, x); y = FA(y, 4.1665795894E-2f);
    y = FM(y, x); y = FA(y, 1.6666665459E-1f);
    y = FM(y, x); y = FA(y, 5.0000001201E-1f);
    y = FM(y, z);
    y = FA(y, x);
    y = FA(y, 1.0f);
    return y;
}
__device__ __noinline__ float cephes_exp_general(float x) {
    if (x != x) return x;
    if (x > 88.72283935546875f) return INFINITY;  // > log(FLT_MAX)
    if (x < -103.98f) return 0.0f;                // < log(2^-150): rounds to +0
    float fx = FM(x, 1.44269504088896341f);       // cephes_LOG2EF
    fx = FA(fx, 0.5f);
    fx = floorf(fx);
    float y = cephes_exp_poly(x, fx);
    const int n = (int)fx;
    if (n > 127) {  // the reference clamps x to 88.376 instead (avx_mathfun.h:360)
        y = FM(y, __uint_as_float((unsigned)(n - 1 + 0x7f) << 23));
        return FM(y, 2.0f);
    }
    if (n < -126) {  // denormal result: scale in two exact / once-rounded steps
        y = FM(y, __uint_as_float((unsigned)(n + 24 + 0x7f) << 23));
        return FM(y, 5.9604644775390625e-08f);  // 2^-24
    }
    return FM(y, __uint_as_float((unsigned)(n + 0x7f) << 23));
}
__device__ __forceinline__ float cephes_expf(float x) {
    if (!(fabsf(x) <= 87.0f)) return cephes_exp_general(x);  // n stays within [-126, 126]
    float v = FM(x, 1.44269504088896341f);
    v = FA(v, 0.5f);
    // floor(v) by a round-down add of 1.5 * 2^23: t = 1.5*2^23 + floor(v) exactly, |v| < 2^22
    const float t = __fadd_rd(v, 12582912.0f);
    const float fx = FS(t, 12582912.0f);
    const float y = cephes_exp_poly(x, fx);
    // bits(t) = 0x4B400000 + floor(v)  ->  2^n = (n + 127) << 23
    const unsigned pow2n = (__float_as_uint(t) - 0x4B400000u + 0x7fu) << 23;
    return FM(y, __uint_as_float(pow2n));
}

// ---- sin / cos -----------------------------------------------------------------------------
// avx_mathfun.h:506-601 (sin256_ps) and :604-673 (cos256_ps).
// `jb` carries j in its low bits (the exponent bits of the magic float above them are harmless:
// only bits 1 and 2 are ever tested).
__device__ __forceinline__ float cephes_sincos_core(float x, float y, bool sin_poly, unsigned sign) {
    // extended precision modular arithmetic: x = ((x - y*DP1) - y*DP2) - y*DP3
    const float x1 = FM(y, -0.78515625f);
    const float x2 = FM(y, -2.4187564849853515625e-4f);
    const float x3 = FM(y, -3.77489497744594108e-8f);
    x = FA(x, x1);
    x = FA(x, x2);
    x = FA(x, x3);
    const float z = FM(x, x);
    // one Horner chain, coefficients of the polynomial this lane needs:
    //   cos: ((c0 z + c1) z + c2) z * z - z/2 + 1      sin: ((s0 z + s1) z + s2) z * x + x
    float t = sin_poly ? -1.9515295891E-4f : 2.443315711809948E-005f;
    t = FM(t, z); t = FA(t, sin_poly ? 8.3321608736E-3f : -1.388731625493765E-003f);
    t = FM(t, z); t = FA(t, sin_poly ? -1.6666654611E-1f : 4.166664568298827E-002f);
    t = FM(t, z);
    t = FM(t, sin_poly ? x : z);
    // cos: t - z*0.5 == t + z*(-0.5) exactly;  the reference's and/andnot/add select then adds +0.0
    // to the chosen polynomial: a no-op for cos (its value is >= 0.7), "+ 0.0" for sin
    t = FA(t, sin_poly ? x : FM(z, -0.5f));
    t = FA(t, sin_poly ? 0.0f : 1.0f);
    return __uint_as_float(__float_as_uint(t) ^ sign);
}
__device__ __noinline__ float sin_general(float x) { return sinf(x); }
__device__ __noinline__ float cos_general(float x) { return cosf(x); }

__device__ __forceinline__ float cephes_sinf(float xin) {
    const float x = fabsf(xin);
    if (!(x <= 8192.0f)) return sin_general(xin);  // outside the documented Cephes range, inf, NaN
    // j = ((int)y + 1) & ~1 ;  y = (float)j      (y < 2^14 here)
    const float yt = __fadd_rz(FM(x, 1.27323954473516f), 8388608.0f);   // 2^23 + trunc(x * 4/pi)
    const unsigned jb = (__float_as_uint(yt) + 1u) & ~1u;
    const float y = FS(__uint_as_float(jb), 8388608.0f);
    const unsigned sign = (__float_as_uint(xin) & 0x80000000u) ^ ((jb & 4u) << 29);
    return cephes_sincos_core(x, y, (jb & 2u) == 0u, sign);
}

__device__ __forceinline__ float cephes_cosf(float xin) {
    const float x = fabsf(xin);
    if (!(x <= 8192.0f)) return cos_general(xin);
    const float yt = __fadd_rz(FM(x, 1.27323954473516f), 8388608.0f);
    unsigned jb = (__float_as_uint(yt) + 1u) & ~1u;
    const float y = FS(__uint_as_float(jb), 8388608.0f);
    jb -= 2u;  // low three bits behave exactly as two's-complement j - 2
    const unsigned sign = (~jb & 4u) << 29;
    return cephes_sincos_core(x, y, (jb & 2u) == 0u, sign);
}

#undef FM
#undef FA
#undef FS

// Loads in flight per thread matter as much as the instruction count: with one 32-byte load per
// thread a warp that is computing has nothing in flight, and 64 warps/SM x 32 B no longer cover the
// HBM latency-bandwidth product (sin: 5.8 TB/s at kUnroll 1, 6.7 at 4, 5.8 again at 8 where the
// registers cut occupancy).  Measured with tools/time_unary.py, 2^28 elements, B200:
//   kUnroll 4: sin 6.72, cos 6.73, exp 6.66, log 6.42 TB/s.
// A packed two-lane variant (PTX fma.rn.f32x2 -> FFMA2, products as fma(a,b,-0.0), sums as
// fma(a,1.0,b) with the constants hidden in __constant__ memory so that ptxas cannot re-fold and
// then CONTRACT the pairs -- it does, even under -fmad=false) was bit-exact but no faster: FFMA2
// retires two lanes at the cost of two scalar instructions on this part.  It was removed.
struct SinF32 { typedef float In; typedef float Out; static constexpr int kUnroll = 4; static __device__ __forceinline__ float apply(float a) { return cephes_sinf(a); } };
struct CosF32 { typedef float In; typedef float Out; static constexpr int kUnroll = 4; static __device__ __forceinline__ float apply(float a) { return cephes_cosf(a); } };
struct LogF32 { typedef float In; typedef float Out; static constexpr int kUnroll = 4; static __device__ __forceinline__ float apply(float a) { return cephes_logf(a); } };
struct ExpF32 { typedef float In; typedef float Out; static constexpr int kUnroll = 4; static __device__ __forceinline__ float apply(float a) { return cephes_expf(a); } };
struct SinF64 { typedef double In; typedef double Out; static constexpr int kUnroll = 2; static __device__ __forceinline__ double apply(double a) { return sin(a); } };
struct CosF64 { typedef double In; typedef double Out; static constexpr int kUnroll = 2; static __device__ __forceinline__ double apply(double a) { return cos(a); } };
// float64 log.  CUDA's log() costs ~50 FP64-pipe operations per element, which makes the kernel
// FP64-bound on B200 (64 lanes/clk/SM): 5.7 TB/s.  This is the classic s = f/(2+f) formulation
// (Sun fdlibm e_log.c: argument reduced to [sqrt(2)/2, sqrt(2)), 7-term minimax polynomial in s^2,
// ln2 split hi/lo) with the division replaced by MUFU.RCP64H + two Newton steps -- `s` only scales
// the O(f^2) correction term, so a reciprocal good to ~1e-16 costs nothing in accuracy.  About 28
// FP64 operations; measured error < 1 ULP (tests: <= 2 ULP vs long double / decimal).
// Zero, negatives, inf, NaN and denormals take CUDA's log() on a __noinline__ path.
__device__ __noinline__ double log_f64_general(double x) { return log(x); }

__device__ __forceinline__ double log_f64(double x) {
    const unsigned long long ix = (unsigned long long)__double_as_longlong(x);
    if (ix - 0x0010000000000000ULL >= 0x7fe0000000000000ULL) return log_f64_general(x);  // not positive normal finite
    int hx = (int)(ix >> 32);
    int k = (hx >> 20) - 1023;
    hx &= 0x000fffff;
    const int i = (hx + 0x95f64) & 0x100000;                // 1 if the mantissa is above sqrt(2)
    const double m = __hiloint2double(hx | (i ^ 0x3ff00000), (int)(unsigned)ix);  // m in [sqrt(2)/2, sqrt(2))
    k += i >> 20;
    const double f = m - 1.0;
    const double d = 2.0 + f;
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));   // ~2^-23 relative
    double e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    e = fma(-d, r, 1.0);
    r = fma(r, e, r);                                       // ~2^-92 before rounding
    const double s = f * r;
    const double dk = (double)k;
    const double z = s * s;
    const double w = z * z;
    const double t1 = w * fma(w, fma(w, 1.531383769920937332e-01, 2.222219843214978396e-01), 3.999999999940941908e-01);
    const double t2 = z * fma(w, fma(w, fma(w, 1.479819860511658591e-01, 1.818357216161805012e-01), 2.857142874366239149e-01),
                              6.666666666666735130e-01);
    const double R = t2 + t1;
    const double hfsq = 0.5 * f * f;
    // k*ln2_hi - ((hfsq - (s*(hfsq+R) + k*ln2_lo)) - f)
    return dk * 6.93147180369123816490e-01 - ((hfsq - (s * (hfsq + R) + dk * 1.90821492927058770002e-10)) - f);
}
struct LogF64 { typedef double In; typedef double Out; static constexpr int kUnroll = 2; static __device__ __forceinline__ double apply(double a) { return log_f64(a); } };
struct ExpF64 { typedef double In; typedef double Out; static constexpr int kUnroll = 2; static __device__ __forceinline__ double apply(double a) { return exp(a); } };

}  // namespace

namespace pncu { pncu_unary_fn trig_more(int func, int t); }

extern "C" pncu_unary_fn pncu_GetTrigOpFast(int func, int t, int* wantedOutType) {
    if (!wantedOutType || func <= PNCU_TRIG_INVALID || func >= PNCU_TRIG_LAST) return NULL;
    // ops_trig.cpp:584-777: every op sets the out-type to the in-type, body or not -- the hook
    // layer registers all 20 names for f32/f64 (src/pnumpy/_pnumpy.cpp:1492-1547).
    *wantedOutType = t;
    switch (func) {
        case PNCU_SIN:
            if (t == PNCU_FLOAT) return launch_unary<SinF32>;
            if (t == PNCU_DOUBLE) return launch_unary<SinF64>;
            break;
        case PNCU_COS:
            if (t == PNCU_FLOAT) return launch_unary<CosF32>;
            if (t == PNCU_DOUBLE) return launch_unary<CosF64>;
            break;
        case PNCU_LOG:
            if (t == PNCU_FLOAT) return launch_unary<LogF32>;
            if (t == PNCU_DOUBLE) return launch_unary<LogF64>;
            break;
        case PNCU_EXP:
            if (t == PNCU_FLOAT) return launch_unary<ExpF32>;
            if (t == PNCU_DOUBLE) return launch_unary<ExpF64>;
            break;
        default:
            // the sixteen ufuncs the reference hooks without a body (ops_trig_more.cu)
            return trig_more(func, t);
    }
    return NULL;
}
