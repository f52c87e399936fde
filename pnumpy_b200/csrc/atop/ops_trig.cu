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
// clamp :326-327, float coefficients :472-481), so there is nothing to match; the kernels use the
// hand-written bodies of f64_math.cuh (<= 1.5 ULP measured; CUDA's functions outside their fast
// domain) and are checked against correctly rounded values.
#include <float.h>
#include <math.h>
#include "ew.cuh"
#include "functors.cuh"
#include "f64_math.cuh"

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
// fast body: x positive, normal, finite.  Same values as cephes_log_core(x, 0) with the exponent /
// mantissa split done in integers: adding 0x3f800000 - bits(SQRTHF) carries into the exponent field
// exactly when the mantissa is >= sqrt(1/2), so E' = E + !lt and the mantissa lands in
// [sqrt(1/2), sqrt(2)) -- and m' - 1 is exactly the reference's (x - 1) + (lt ? x : 0), both being exact.
__device__ __forceinline__ float cephes_log_fast(float x) {
    const unsigned ix = __float_as_uint(x) + 0x004afb0du;                 // 0x3f800000 - 0x3f3504f3
    const float e = FS(__uint_as_float(0x4B000000u + (ix >> 23)), 8388735.0f);  // (2^23 + E') - (2^23 + 127)
    x = FS(__uint_as_float((ix & 0x007fffffu) + 0x3f3504f3u), 1.0f);
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
__device__ __forceinline__ float cephes_logf(float x) {
    // positive normal finite <=> bits in [0x00800000, 0x7f800000)
    if (__float_as_uint(x) - 0x00800000u >= 0x7f000000u) return cephes_log_general(x);
    return cephes_log_fast(x);
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
__device__ __forceinline__ float cephes_exp_fast(float x) {  // |x| <= 87: n stays within [-126, 126]
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
__device__ __forceinline__ float cephes_expf(float x) {
    if (!(fabsf(x) <= 87.0f)) return cephes_exp_general(x);
    return cephes_exp_fast(x);
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

__device__ __forceinline__ float cephes_sinf_fast(float xin) {  // |x| <= 8192
    const float x = fabsf(xin);
    // j = ((int)y + 1) & ~1 ;  y = (float)j      (y < 2^14 here)
    const float yt = __fadd_rz(FM(x, 1.27323954473516f), 8388608.0f);   // 2^23 + trunc(x * 4/pi)
    const unsigned jb = (__float_as_uint(yt) + 1u) & ~1u;
    const float y = FS(__uint_as_float(jb), 8388608.0f);
    const unsigned sign = (__float_as_uint(xin) & 0x80000000u) ^ ((jb & 4u) << 29);
    return cephes_sincos_core(x, y, (jb & 2u) == 0u, sign);
}
__device__ __forceinline__ float cephes_sinf(float xin) {
    if (!(fabsf(xin) <= 8192.0f)) return sin_general(xin);  // outside the documented Cephes range, inf, NaN
    return cephes_sinf_fast(xin);
}

__device__ __forceinline__ float cephes_cosf_fast(float xin) {
    const float x = fabsf(xin);
    const float yt = __fadd_rz(FM(x, 1.27323954473516f), 8388608.0f);
    unsigned jb = (__float_as_uint(yt) + 1u) & ~1u;
    const float y = FS(__uint_as_float(jb), 8388608.0f);
    jb -= 2u;  // low three bits behave exactly as two's-complement j - 2
    const unsigned sign = (~jb & 4u) << 29;
    return cephes_sincos_core(x, y, (jb & 2u) == 0u, sign);
}
__device__ __forceinline__ float cephes_cosf(float xin) {
    if (!(fabsf(xin) <= 8192.0f)) return cos_general(xin);
    return cephes_cosf_fast(xin);
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
//
// Every functor below has a branch-free fast body (`fast`, valid where `ok`) and an out-of-line
// general function (`apply` = ok ? fast : IEEE edge cases), see ew.cuh: the kernel tests a thread's
// whole set of inputs once, so a value's result never depends on its neighbours, only the speed does.
__device__ __noinline__ float sinf_elem(float a) { return cephes_sinf(a); }
__device__ __noinline__ float cosf_elem(float a) { return cephes_cosf(a); }
__device__ __noinline__ float logf_elem(float a) { return cephes_logf(a); }
__device__ __noinline__ float expf_elem(float a) { return cephes_expf(a); }

struct SinF32 {
    typedef float In; typedef float Out;
    static constexpr int kUnroll = 4; static constexpr bool kHasFast = true;
    static __device__ __forceinline__ bool ok(float a) { return fabsf(a) <= 8192.0f; }
    static __device__ __forceinline__ float fast(float a) { return cephes_sinf_fast(a); }
    static __device__ __forceinline__ float apply(float a) { return sinf_elem(a); }
};
struct CosF32 {
    typedef float In; typedef float Out;
    static constexpr int kUnroll = 4; static constexpr bool kHasFast = true;
    static __device__ __forceinline__ bool ok(float a) { return fabsf(a) <= 8192.0f; }
    static __device__ __forceinline__ float fast(float a) { return cephes_cosf_fast(a); }
    static __device__ __forceinline__ float apply(float a) { return cosf_elem(a); }
};
struct LogF32 {
    typedef float In; typedef float Out;
    static constexpr int kUnroll = 4; static constexpr bool kHasFast = true;
    static __device__ __forceinline__ bool ok(float a) { return __float_as_uint(a) - 0x00800000u < 0x7f000000u; }
    static __device__ __forceinline__ float fast(float a) { return cephes_log_fast(a); }
    static __device__ __forceinline__ float apply(float a) { return logf_elem(a); }
};
struct ExpF32 {
    typedef float In; typedef float Out;
    static constexpr int kUnroll = 4; static constexpr bool kHasFast = true;
    static __device__ __forceinline__ bool ok(float a) { return fabsf(a) <= 87.0f; }
    static __device__ __forceinline__ float fast(float a) { return cephes_exp_fast(a); }
    static __device__ __forceinline__ float apply(float a) { return expf_elem(a); }
};

// ---- float64 -------------------------------------------------------------------------------------
// The reference's pd bodies are defective, so there is nothing to be bit-identical
// to; the target is <= 2 ULP of the correctly rounded value.  Bodies: f64_math.cuh (why they are
// hand-written, their cost and their measured accuracy are at its top).
__device__ __noinline__ double sin64_elem(double a) { return f64_sincos_ok(a) ? f64_sincos_fast<0>(a) : sin(a); }
__device__ __noinline__ double cos64_elem(double a) { return f64_sincos_ok(a) ? f64_sincos_fast<1>(a) : cos(a); }
__device__ __noinline__ double exp64_elem(double a) { return f64_exp_ok(a) ? f64_exp_fast(a) : exp(a); }
__device__ __noinline__ double log64_elem(double a) {
    if (!f64_log_ok(a)) return log(a);
    const int i = f64_log_index(a);
    return f64_log_fast(a, kLogTab[2 * i], kLogTab[2 * i + 1]);  // flat table in constant memory: the rare path
}

struct SinF64 {
    typedef double In; typedef double Out;
    static constexpr int kUnroll = 2; static constexpr bool kHasFast = true;
    static __device__ __forceinline__ bool ok(double a) { return f64_sincos_ok(a); }
    static __device__ __forceinline__ double fast(double a) { return f64_sincos_fast<0>(a); }
    static __device__ __forceinline__ double apply(double a) { return sin64_elem(a); }
};
struct CosF64 {
    typedef double In; typedef double Out;
    static constexpr int kUnroll = 2; static constexpr bool kHasFast = true;
    static __device__ __forceinline__ bool ok(double a) { return f64_sincos_ok(a); }
    static __device__ __forceinline__ double fast(double a) { return f64_sincos_fast<1>(a); }
    static __device__ __forceinline__ double apply(double a) { return cos64_elem(a); }
};
struct ExpF64 {
    typedef double In; typedef double Out;
    static constexpr int kUnroll = 2; static constexpr bool kHasFast = true;
    static __device__ __forceinline__ bool ok(double a) { return f64_exp_ok(a); }
    static __device__ __forceinline__ double fast(double a) { return f64_exp_fast(a); }
    static __device__ __forceinline__ double apply(double a) { return exp64_elem(a); }
};
// The log table once more in plain global memory: building the shared copy reads it with ordinary
// (L2-resident, warp-coalesced) loads; a __constant__ array would serialise the 4 different entries a
// warp reads per step.
static __device__ const unsigned long long gLogTab[256] = F64_LOGTAB_INIT;

// log: the 128-entry table lives in shared memory, 8 copies interleaved at 16-byte granularity:
// copy j of entry i is at byte (i*8 + j)*16, and lane l reads copy l % 8, so the 8 lanes of a quarter
// warp (one wavefront of an LDS.128) hit 8 different groups of 4 banks whatever their indices are (ncu: 2 % extra
// wavefronts from bank conflicts on random data, profiles/r01_trig_full.md; unroll 4 and a 64-register cap were tried:
// no change, the kernel sits at 82 % of DRAM throughput with 51 % issue and 39 % FP64-pipe utilisation).
struct LogF64 {
    typedef double In; typedef double Out;
    static constexpr int kUnroll = 2; static constexpr bool kHasFast = true;
    static constexpr int kSmemBytes = 128 * 8 * 16;
    static __device__ __forceinline__ void smem_init(void* smem) {
        ulonglong2* t = reinterpret_cast<ulonglong2*>(smem);
        const ulonglong2* g = reinterpret_cast<const ulonglong2*>(gLogTab);
        // slot s = 8*entry + copy: the 8 lanes of a quarter warp write 8 different bank groups
#pragma unroll
        for (int s = threadIdx.x; s < 128 * 8; s += EW_BLOCK) t[s] = __ldg(g + (s >> 3));
    }
    static __device__ __forceinline__ const void* smem_lane(const void* smem) {
        return reinterpret_cast<const unsigned char*>(smem) + (threadIdx.x & 7) * 16;
    }
    static __device__ __forceinline__ bool ok(double a) { return f64_log_ok(a); }
    static __device__ __forceinline__ double fast(double a, const void* table) {
        const ulonglong2 e = *reinterpret_cast<const ulonglong2*>(reinterpret_cast<const unsigned char*>(table) + f64_log_index(a) * 128);
        return f64_log_fast(a, e.x, e.y);
    }
    static __device__ __forceinline__ double apply(double a) { return log64_elem(a); }
};

// log2 / log10 / log1p / expm1 (float64): the same machinery with a different recombination
// (f64_math.cuh); float32 and the other twelve ufuncs of SURVEY.md 8(f) f-4 are in ops_trig_more.cu.
__device__ __noinline__ double log2_64_elem(double a) {
    if (!f64_log_ok(a)) return log2(a);
    const int i = f64_log_index(a);
    return f64_log2_fast(a, kLogTab[2 * i], kLogTab[2 * i + 1]);
}
__device__ __noinline__ double log10_64_elem(double a) {
    if (!f64_log_ok(a)) return log10(a);
    const int i = f64_log_index(a);
    return f64_log10_fast(a, kLogTab[2 * i], kLogTab[2 * i + 1]);
}
__device__ __noinline__ double log1p_64_elem(double a) {
    if (!f64_log1p_ok(a)) return log1p(a);
    const int i = f64_log1p_index(a);
    return f64_log1p_fast(a, kLogTab[2 * i], kLogTab[2 * i + 1]);
}
__device__ __noinline__ double expm1_64_elem(double a) { return f64_exp_ok(a) ? f64_expm1_fast(a) : expm1(a); }

#define PNCU_LOG_TABLE_FUNCTOR(NAME, OK, INDEX, FAST, ELEM)                                                        \
    struct NAME {                                                                                                  \
        typedef double In; typedef double Out;                                                                     \
        static constexpr int kUnroll = 2; static constexpr bool kHasFast = true;                                   \
        static constexpr int kSmemBytes = LogF64::kSmemBytes;                                                      \
        static __device__ __forceinline__ void smem_init(void* smem) { LogF64::smem_init(smem); }                  \
        static __device__ __forceinline__ const void* smem_lane(const void* smem) { return LogF64::smem_lane(smem); } \
        static __device__ __forceinline__ bool ok(double a) { return OK(a); }                                      \
        static __device__ __forceinline__ double fast(double a, const void* table) {                               \
            const ulonglong2 e = *reinterpret_cast<const ulonglong2*>(reinterpret_cast<const unsigned char*>(table) + INDEX(a) * 128); \
            return FAST(a, e.x, e.y);                                                                              \
        }                                                                                                          \
        static __device__ __forceinline__ double apply(double a) { return ELEM(a); }                               \
    };
// atanh / asinh / acosh: log1p of an argument computed in the body, so they take the table accessor
struct SmemLogTable {  // this lane's copy of the bank-replicated shared-memory table (LogF64::smem_lane)
    const void* base;
    __device__ __forceinline__ F64LogEntry get(int i) const {
        const ulonglong2 e = *reinterpret_cast<const ulonglong2*>(reinterpret_cast<const unsigned char*>(base) + i * 128);
        F64LogEntry r; r.e0 = e.x; r.e1 = e.y; return r;
    }
};
__device__ __noinline__ double atanh64_elem(double a) { return f64_atanh_ok(a) ? f64_atanh_fast(a, F64FlatLogTable()) : atanh(a); }
__device__ __noinline__ double asinh64_elem(double a) { return f64_asinh_ok(a) ? f64_asinh_fast(a, F64FlatLogTable()) : asinh(a); }
__device__ __noinline__ double acosh64_elem(double a) { return f64_acosh_ok(a) ? f64_acosh_fast(a, F64FlatLogTable()) : acosh(a); }
#define PNCU_LOG_ACCESSOR_FUNCTOR(NAME, OK, FAST, ELEM)                                                            \
    struct NAME {                                                                                                  \
        typedef double In; typedef double Out;                                                                     \
        static constexpr int kUnroll = 2; static constexpr bool kHasFast = true;                                   \
        static constexpr int kSmemBytes = LogF64::kSmemBytes;                                                      \
        static __device__ __forceinline__ void smem_init(void* smem) { LogF64::smem_init(smem); }                  \
        static __device__ __forceinline__ const void* smem_lane(const void* smem) { return LogF64::smem_lane(smem); } \
        static __device__ __forceinline__ bool ok(double a) { return OK(a); }                                      \
        static __device__ __forceinline__ double fast(double a, const void* table) { return FAST(a, SmemLogTable{table}); } \
        static __device__ __forceinline__ double apply(double a) { return ELEM(a); }                               \
    };
PNCU_LOG_ACCESSOR_FUNCTOR(AtanhF64, f64_atanh_ok, f64_atanh_fast, atanh64_elem)
PNCU_LOG_ACCESSOR_FUNCTOR(AsinhF64, f64_asinh_ok, f64_asinh_fast, asinh64_elem)
PNCU_LOG_ACCESSOR_FUNCTOR(AcoshF64, f64_acosh_ok, f64_acosh_fast, acosh64_elem)
#undef PNCU_LOG_ACCESSOR_FUNCTOR
PNCU_LOG_TABLE_FUNCTOR(Log2F64, f64_log_ok, f64_log_index, f64_log2_fast, log2_64_elem)
PNCU_LOG_TABLE_FUNCTOR(Log10F64, f64_log_ok, f64_log_index, f64_log10_fast, log10_64_elem)
PNCU_LOG_TABLE_FUNCTOR(Log1pF64, f64_log1p_ok, f64_log1p_index, f64_log1p_fast, log1p_64_elem)
#undef PNCU_LOG_TABLE_FUNCTOR
__device__ __noinline__ double tanh64_elem(double a) { return f64_tanh_ok(a) ? f64_tanh_fast(a) : tanh(a); }
__device__ __noinline__ double sinh64_elem(double a) { return f64_exp_ok(a) ? f64_sinh_fast(a) : sinh(a); }
__device__ __noinline__ double tan64_elem(double a) { return f64_sincos_ok(a) ? f64_tan_fast(a) : tan(a); }
#define PNCU_F64_FAST_FUNCTOR(NAME, OK, FAST, ELEM) PNCU_F64_FAST_FUNCTOR_U(NAME, OK, FAST, ELEM, 2)
#define PNCU_F64_FAST_FUNCTOR_U(NAME, OK, FAST, ELEM, UNROLL)                                      \
    struct NAME {                                                                                  \
        typedef double In; typedef double Out;                                                     \
        static constexpr int kUnroll = UNROLL; static constexpr bool kHasFast = true;              \
        static __device__ __forceinline__ bool ok(double a) { return OK(a); }                      \
        static __device__ __forceinline__ double fast(double a) { return FAST(a); }                \
        static __device__ __forceinline__ double apply(double a) { return ELEM(a); }               \
    };
__device__ __noinline__ double atan64_elem(double a) { return f64_atan_ok(a) ? f64_atan_fast(a) : atan(a); }
__device__ __noinline__ double asin64_elem(double a) { return f64_asin_ok(a) ? f64_asin_fast(a) : asin(a); }
__device__ __noinline__ double acos64_elem(double a) { return f64_asin_ok(a) ? f64_acos_fast(a) : acos(a); }
PNCU_F64_FAST_FUNCTOR(AtanF64, f64_atan_ok, f64_atan_fast, atan64_elem)
PNCU_F64_FAST_FUNCTOR(AsinF64, f64_asin_ok, f64_asin_fast, asin64_elem)
PNCU_F64_FAST_FUNCTOR(AcosF64, f64_asin_ok, f64_acos_fast, acos64_elem)
PNCU_F64_FAST_FUNCTOR(TanhF64, f64_tanh_ok, f64_tanh_fast, tanh64_elem)
PNCU_F64_FAST_FUNCTOR(SinhF64, f64_exp_ok, f64_sinh_fast, sinh64_elem)
PNCU_F64_FAST_FUNCTOR(TanF64, f64_sincos_ok, f64_tan_fast, tan64_elem)
#undef PNCU_F64_FAST_FUNCTOR
#undef PNCU_F64_FAST_FUNCTOR_U
__device__ __noinline__ double exp2_64_elem(double a) { return f64_exp2_ok(a) ? f64_exp2_fast(a) : exp2(a); }
struct Exp2F64 {
    typedef double In; typedef double Out;
    static constexpr int kUnroll = 2; static constexpr bool kHasFast = true;
    static __device__ __forceinline__ bool ok(double a) { return f64_exp2_ok(a); }
    static __device__ __forceinline__ double fast(double a) { return f64_exp2_fast(a); }
    static __device__ __forceinline__ double apply(double a) { return exp2_64_elem(a); }
};
struct Expm1F64 {
    typedef double In; typedef double Out;
    static constexpr int kUnroll = 2; static constexpr bool kHasFast = true;
    static __device__ __forceinline__ bool ok(double a) { return f64_exp_ok(a); }
    static __device__ __forceinline__ double fast(double a) { return f64_expm1_fast(a); }
    static __device__ __forceinline__ double apply(double a) { return expm1_64_elem(a); }
};

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
        case PNCU_LOG2:
            if (t == PNCU_DOUBLE) return launch_unary<Log2F64>;
            return trig_more(func, t);
        case PNCU_LOG10:
            if (t == PNCU_DOUBLE) return launch_unary<Log10F64>;
            return trig_more(func, t);
        case PNCU_LOG1P:
            if (t == PNCU_DOUBLE) return launch_unary<Log1pF64>;
            return trig_more(func, t);
        case PNCU_EXPM1:
            if (t == PNCU_DOUBLE) return launch_unary<Expm1F64>;
            return trig_more(func, t);
        case PNCU_EXP2:
            if (t == PNCU_DOUBLE) return launch_unary<Exp2F64>;
            return trig_more(func, t);
        case PNCU_ATAN:
            if (t == PNCU_DOUBLE) return launch_unary<AtanF64>;
            return trig_more(func, t);
        case PNCU_ASIN:
            if (t == PNCU_DOUBLE) return launch_unary<AsinF64>;
            return trig_more(func, t);
        case PNCU_ACOS:
            if (t == PNCU_DOUBLE) return launch_unary<AcosF64>;
            return trig_more(func, t);
        case PNCU_ATANH:
            if (t == PNCU_DOUBLE) return launch_unary<AtanhF64>;
            return trig_more(func, t);
        case PNCU_ASINH:
            if (t == PNCU_DOUBLE) return launch_unary<AsinhF64>;
            return trig_more(func, t);
        case PNCU_ACOSH:
            if (t == PNCU_DOUBLE) return launch_unary<AcoshF64>;
            return trig_more(func, t);
        case PNCU_TANH:
            if (t == PNCU_DOUBLE) return launch_unary<TanhF64>;
            return trig_more(func, t);
        case PNCU_SINH:
            if (t == PNCU_DOUBLE) return launch_unary<SinhF64>;
            return trig_more(func, t);
        case PNCU_TAN:
            if (t == PNCU_DOUBLE) return launch_unary<TanF64>;
            return trig_more(func, t);
        default:
            // the other ufuncs the reference hooks without a body (ops_trig_more.cu)
            return trig_more(func, t);
    }
    return NULL;
}
