// The sixteen transcendental ufuncs the reference hooks WITHOUT a body of its own -- tan, arcsin,
// arccos, arctan, sinh, cosh, tanh, arcsinh, arccosh, arctanh, cbrt, exp2, expm1, log2, log10, log1p
// (GetTrigOpFast returns NULL for them, src/atop/ops_trig.cpp:599-777, and the dispatcher threads
// NumPy's own loop over 16K chunks, src/pnumpy/_pnumpy.cpp:973-993).  SURVEY.md 8(f) row f-4.
//
// Here they get a GPU body.  This file holds the float32 ones and float64 cosh, cbrt, inside the same
// streaming kernel as sin/log/exp (ew1_vec_kernel, several 256-bit loads in flight per thread).  Where CUDA's
// math library streams at the HBM rate it is used as is (6.8 TB/s: profiles/r01_transcendentals.md); float32
// asinh / acosh / atanh did not (4.2 / 5.6 / 6.3 TB/s, issue-bound) and have lean branch-free bodies in
// f32_math.cuh.  The other fourteen float64 ones have lean bodies in f64_math.cuh and are picked in ops_trig.cu
// before this getter is asked.  There is no reference arithmetic
// to be bit-identical to -- NumPy's loops call the platform libm / SVML, whose results differ from
// one libm to the next in the last place -- so the parity bar is stated in ULP against the
// correctly rounded value: <= 4 ULP (tests/test_gpu_parity.py; CUDA documents 1-4 ULP for these),
// with the IEEE special cases (NaN, +-inf, +-0, domain errors -> NaN, poles -> +-inf) exact.
#include "ew.cuh"
#include "functors.cuh"
#include "f32_math.cuh"

using namespace pncu;

namespace {

#define PNCU_LIBM_F32(NAME, FN32)                                                                              \
    struct NAME##F32 { typedef float In; typedef float Out; static constexpr int kUnroll = 4;                 \
                       static __device__ __forceinline__ float apply(float a) { return FN32(a); } };
#define PNCU_LIBM_F64(NAME, FN64)                                                                              \
    struct NAME##F64 { typedef double In; typedef double Out; static constexpr int kUnroll = 1;               \
                       static __device__ __forceinline__ double apply(double a) { return FN64(a); } };

PNCU_LIBM_F32(Tan, tanf)
PNCU_LIBM_F32(Asin, asinf)
PNCU_LIBM_F32(Acos, acosf)
PNCU_LIBM_F32(Atan, atanf)
PNCU_LIBM_F32(Sinh, sinhf)
PNCU_LIBM_F32(Cosh, coshf)
PNCU_LIBM_F32(Tanh, tanhf)
PNCU_LIBM_F32(Log2, log2f)
PNCU_LIBM_F32(Log10, log10f)
PNCU_LIBM_F32(Exp2, exp2f)
PNCU_LIBM_F32(Expm1, expm1f)
PNCU_LIBM_F32(Log1p, log1pf)
PNCU_LIBM_F32(Cbrt, cbrtf)
PNCU_LIBM_F64(Cosh, cosh)
PNCU_LIBM_F64(Cbrt, cbrt)
#undef PNCU_LIBM_F32
#undef PNCU_LIBM_F64

// fast body on the fast domain (tested once per thread, ew.cuh), CUDA's function out of line elsewhere
__device__ __noinline__ float asinhf_elem(float a) { return f32_asinh_ok(a) ? f32_asinh_fast(a) : asinhf(a); }
__device__ __noinline__ float acoshf_elem(float a) { return f32_acosh_ok(a) ? f32_acosh_fast(a) : acoshf(a); }
__device__ __noinline__ float atanhf_elem(float a) { return f32_atanh_ok(a) ? f32_atanh_fast(a) : atanhf(a); }
#define PNCU_F32_FAST_FUNCTOR(NAME, OK, FAST, ELEM)                                                \
    struct NAME {                                                                                  \
        typedef float In; typedef float Out;                                                       \
        static constexpr int kUnroll = 4; static constexpr bool kHasFast = true;                   \
        static __device__ __forceinline__ bool ok(float a) { return OK(a); }                       \
        static __device__ __forceinline__ float fast(float a) { return FAST(a); }                  \
        static __device__ __forceinline__ float apply(float a) { return ELEM(a); }                 \
    };
PNCU_F32_FAST_FUNCTOR(AsinhF32, f32_asinh_ok, f32_asinh_fast, asinhf_elem)
PNCU_F32_FAST_FUNCTOR(AcoshF32, f32_acosh_ok, f32_acosh_fast, acoshf_elem)
PNCU_F32_FAST_FUNCTOR(AtanhF32, f32_atanh_ok, f32_atanh_fast, atanhf_elem)
#undef PNCU_F32_FAST_FUNCTOR

}  // namespace

namespace pncu {

pncu_unary_fn trig_more(int func, int t) {
#define PNCU_PICK(OP, NAME)                                        \
    case OP:                                                       \
        if (t == PNCU_FLOAT) return launch_unary<NAME##F32>;       \
        return NULL;
#define PNCU_PICK2(OP, NAME)                                       \
    case OP:                                                       \
        if (t == PNCU_FLOAT) return launch_unary<NAME##F32>;       \
        if (t == PNCU_DOUBLE) return launch_unary<NAME##F64>;      \
        return NULL;
    switch (func) {
        PNCU_PICK(PNCU_TAN, Tan)
        PNCU_PICK(PNCU_ASIN, Asin)
        PNCU_PICK(PNCU_ACOS, Acos)
        PNCU_PICK(PNCU_ATAN, Atan)
        PNCU_PICK(PNCU_SINH, Sinh)
        PNCU_PICK2(PNCU_COSH, Cosh)
        PNCU_PICK(PNCU_TANH, Tanh)
        PNCU_PICK(PNCU_ASINH, Asinh)
        PNCU_PICK(PNCU_ACOSH, Acosh)
        PNCU_PICK(PNCU_ATANH, Atanh)
        PNCU_PICK(PNCU_LOG2, Log2)
        PNCU_PICK(PNCU_LOG10, Log10)
        PNCU_PICK(PNCU_EXP2, Exp2)
        PNCU_PICK(PNCU_EXPM1, Expm1)
        PNCU_PICK(PNCU_LOG1P, Log1p)
        PNCU_PICK2(PNCU_CBRT, Cbrt)
    }
#undef PNCU_PICK
#undef PNCU_PICK2
    return NULL;
}

}  // namespace pncu
