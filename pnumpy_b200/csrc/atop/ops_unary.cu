// Unary getters: pncu_GetUnaryOpFast (the reference's Fast-then-Slow lookup folded into one).
// replaces: GetUnaryOpFast / GetUnaryOpSlow, src/atop/ops_unary.cpp:1250-1774, and the loop
// bodies UnaryOpFast (:327-376), UnaryNanFast* (:435-928) and the scalar fallbacks (:934-1204).
//
// Semantics follow NumPy's loops (which the reference equals on these ops, SURVEY.md
// App. C.1) with the one documented exception of float `negative`: the reference computes
// 0 - x (src/atop/ops_unary.cpp:295-305) so -(+0.0) = +0.0; NumPy flips the sign bit, as here.
#include "ew.cuh"
#include "functors.cuh"

using namespace pncu;

namespace {

#define PNCU_UNARY_FUNCTOR(NAME, OUT_T, EXPR)                              \
    template <class T> struct NAME {                                       \
        typedef T In;                                                      \
        typedef OUT_T Out;                                                 \
        static __device__ __forceinline__ OUT_T apply(T a) { return EXPR; } \
    };

template <class T> __device__ __forceinline__ T u_abs(T a) {
    typedef typename std::make_unsigned<T>::type U;
    if constexpr (std::is_signed<T>::value) return a < 0 ? (T)(U)((U)0 - (U)a) : a;  // INT_MIN stays INT_MIN
    else return a;
}
template <> __device__ __forceinline__ float u_abs<float>(float a) { return fabsf(a); }
template <> __device__ __forceinline__ double u_abs<double>(double a) { return fabs(a); }

template <class T> __device__ __forceinline__ T u_neg(T a) {
    typedef typename std::make_unsigned<T>::type U;
    return (T)(U)((U)0 - (U)a);
}
template <> __device__ __forceinline__ float u_neg<float>(float a) { return -a; }
template <> __device__ __forceinline__ double u_neg<double>(double a) { return -a; }

// NumPy sign: x>0 -> 1, x<0 -> -1, x==0 -> +0, NaN -> NaN
template <class T> __device__ __forceinline__ T u_sign(T a) {
    return a > 0 ? (T)1 : (a < 0 ? (T)-1 : (a == 0 ? (T)0 : a));
}

__device__ __forceinline__ float u_floor(float a) { return floorf(a); }
__device__ __forceinline__ double u_floor(double a) { return floor(a); }
__device__ __forceinline__ float u_ceil(float a) { return ceilf(a); }
__device__ __forceinline__ double u_ceil(double a) { return ceil(a); }
__device__ __forceinline__ float u_trunc(float a) { return truncf(a); }
__device__ __forceinline__ double u_trunc(double a) { return trunc(a); }
__device__ __forceinline__ float u_rint(float a) { return rintf(a); }  // half-to-even
__device__ __forceinline__ double u_rint(double a) { return rint(a); }
__device__ __forceinline__ float u_sqrt(float a) { return __fsqrt_rn(a); }
__device__ __forceinline__ double u_sqrt(double a) { return __dsqrt_rn(a); }

template <class T> __device__ __forceinline__ bool u_isnan(T a) {
    if constexpr (is_fp<T>::value) return a != a;
    else return false;
}
template <class T> __device__ __forceinline__ bool u_isinf(T a) {
    if constexpr (is_fp<T>::value) return u_abs<T>(a) == (T)INFINITY;
    else return false;
}
template <class T> __device__ __forceinline__ bool u_isfinite(T a) {
    if constexpr (is_fp<T>::value) return u_abs<T>(a) < (T)INFINITY;  // false for inf and NaN
    else return true;
}
__device__ __forceinline__ bool u_signbit(float a) { return (__float_as_uint(a) >> 31) != 0; }
__device__ __forceinline__ bool u_signbit(double a) { return (__double_as_longlong(a) < 0); }
__device__ __forceinline__ bool u_isnormal(float a) {
    const unsigned e = (__float_as_uint(a) >> 23) & 0xffu;
    return e != 0 && e != 0xffu;
}
__device__ __forceinline__ bool u_isnormal(double a) {
    const unsigned e = (unsigned)((unsigned long long)__double_as_longlong(a) >> 52) & 0x7ffu;
    return e != 0 && e != 0x7ffu;
}
template <class T> __device__ __forceinline__ bool u_isnormal(T a) { return a != 0; }

PNCU_UNARY_FUNCTOR(AbsF, T, u_abs<T>(a))
PNCU_UNARY_FUNCTOR(IdentF, T, a)
PNCU_UNARY_FUNCTOR(NegF, T, u_neg<T>(a))
PNCU_UNARY_FUNCTOR(SignF, T, u_sign<T>(a))
PNCU_UNARY_FUNCTOR(InvertF, T, (T)~a)
PNCU_UNARY_FUNCTOR(FloorF, T, u_floor(a))
PNCU_UNARY_FUNCTOR(CeilF, T, u_ceil(a))
PNCU_UNARY_FUNCTOR(TruncF, T, u_trunc(a))
PNCU_UNARY_FUNCTOR(RintF, T, u_rint(a))
PNCU_UNARY_FUNCTOR(SqrtF, T, u_sqrt(a))
PNCU_UNARY_FUNCTOR(LogicalNotF, bool8, (bool8)(a == 0))
PNCU_UNARY_FUNCTOR(IsNanF, bool8, (bool8)u_isnan<T>(a))
PNCU_UNARY_FUNCTOR(IsNotNanF, bool8, (bool8)!u_isnan<T>(a))
PNCU_UNARY_FUNCTOR(IsInfF, bool8, (bool8)u_isinf<T>(a))
PNCU_UNARY_FUNCTOR(IsNotInfF, bool8, (bool8)!u_isinf<T>(a))
PNCU_UNARY_FUNCTOR(IsFiniteF, bool8, (bool8)u_isfinite<T>(a))
PNCU_UNARY_FUNCTOR(IsNotFiniteF, bool8, (bool8)!u_isfinite<T>(a))
PNCU_UNARY_FUNCTOR(IsNormalF, bool8, (bool8)u_isnormal(a))
PNCU_UNARY_FUNCTOR(IsNotNormalF, bool8, (bool8)!u_isnormal(a))
PNCU_UNARY_FUNCTOR(IsNanOrZeroF, bool8, (bool8)(u_isnan<T>(a) || a == 0))
PNCU_UNARY_FUNCTOR(SignBitF, bool8, (bool8)u_signbit(a))

}  // namespace
namespace pncu {
// word-at-a-time forms of the 1- and 2-byte unary ops (see functors.cuh)
#define PNCU_WORD1(FUNCTOR, EXPR)                                                          \
    template <> struct WordOp1<FUNCTOR> {                                                  \
        static constexpr bool enabled = true;                                              \
        static __device__ __forceinline__ unsigned word(unsigned a) { return EXPR; }       \
    };
PNCU_WORD1(AbsF<int8_t>, __vabs4(a)) PNCU_WORD1(AbsF<int16_t>, __vabs2(a))
PNCU_WORD1(NegF<uint8_t>, __vneg4(a)) PNCU_WORD1(NegF<uint16_t>, __vneg2(a))
PNCU_WORD1(InvertF<uint8_t>, ~a) PNCU_WORD1(InvertF<uint16_t>, ~a)
PNCU_WORD1(IdentF<uint8_t>, a) PNCU_WORD1(IdentF<uint16_t>, a)
PNCU_WORD1(LogicalNotF<uint8_t>, __vseteq4(a, 0u)) PNCU_WORD1(LogicalNotF<int8_t>, __vseteq4(a, 0u))
#undef PNCU_WORD1
}  // namespace pncu
namespace {

struct BoolNotF { typedef bool8 In; typedef bool8 Out; static __device__ __forceinline__ bool8 apply(bool8 a) { return (bool8)(a == 0); } };

// Classifiers of an integer input are constants (no NaN/inf among integers): the reference memsets
// (src/atop/ops_unary.cpp:1001-1013, 1651-1659).  A contiguous output is one cudaMemsetAsync -- a
// write-only stream needs no read of the input; anything else takes the generic strided kernel.
template <class T, int V> struct ConstF { typedef T In; typedef bool8 Out; static __device__ __forceinline__ bool8 apply(T) { return (bool8)V; } };
template <class T, int V>
int launch_const(const void* in, void* out, int64_t len, int64_t s1, int64_t so, pncu_stream_t stream) {
    if (len > 0 && in && out && so == 1 && s1 % (int64_t)sizeof(T) == 0) {
        int rc = ensure_init();
        if (rc) return rc;
        if ((rc = check_alias(in, s1, sizeof(T), out, so, 1, len))) return rc;
        PNCU_CUDA(cudaMemsetAsync(out, V, (size_t)len, as_stream(stream)));  // a driver memset, not one of our kernels: not counted
        return 0;
    }
    return launch_unary<ConstF<T, V>>(in, out, len, s1, so, stream);
}
template <int V>
pncu_unary_fn pick_const_int(int t) {
    switch (t) {
        case PNCU_BOOL: case PNCU_UINT8: case PNCU_INT8: return launch_const<uint8_t, V>;
        case PNCU_INT16: case PNCU_UINT16: return launch_const<uint16_t, V>;
        case PNCU_INT32: case PNCU_UINT32: return launch_const<uint32_t, V>;
        case PNCU_INT64: case PNCU_UINT64: return launch_const<uint64_t, V>;
    }
    return NULL;
}
// classifier getter: floats run the predicate kernel, integers the constant
template <template <class> class F, int INT_VALUE>
pncu_unary_fn pick_classifier(int t) {
    if (t == PNCU_FLOAT) return launch_unary<F<float>>;
    if (t == PNCU_DOUBLE) return launch_unary<F<double>>;
    return pick_const_int<INT_VALUE>(t);
}

template <template <class> class F>
pncu_unary_fn pick_all(int t) {
    switch (t) {
        case PNCU_BOOL: case PNCU_UINT8: return launch_unary<F<uint8_t>>;
        case PNCU_INT8: return launch_unary<F<int8_t>>;
        case PNCU_INT16: return launch_unary<F<int16_t>>;
        case PNCU_UINT16: return launch_unary<F<uint16_t>>;
        case PNCU_INT32: return launch_unary<F<int32_t>>;
        case PNCU_UINT32: return launch_unary<F<uint32_t>>;
        case PNCU_INT64: return launch_unary<F<int64_t>>;
        case PNCU_UINT64: return launch_unary<F<uint64_t>>;
        case PNCU_FLOAT: return launch_unary<F<float>>;
        case PNCU_DOUBLE: return launch_unary<F<double>>;
    }
    return NULL;
}
template <template <class> class F>
pncu_unary_fn pick_signed_fp(int t) {
    switch (t) {
        case PNCU_INT8: return launch_unary<F<int8_t>>;
        case PNCU_INT16: return launch_unary<F<int16_t>>;
        case PNCU_INT32: return launch_unary<F<int32_t>>;
        case PNCU_INT64: return launch_unary<F<int64_t>>;
        case PNCU_FLOAT: return launch_unary<F<float>>;
        case PNCU_DOUBLE: return launch_unary<F<double>>;
    }
    return NULL;
}
template <template <class> class F>
pncu_unary_fn pick_fp(int t) {
    switch (t) {
        case PNCU_FLOAT: return launch_unary<F<float>>;
        case PNCU_DOUBLE: return launch_unary<F<double>>;
    }
    return NULL;
}
template <template <class> class F>
pncu_unary_fn pick_uint_width(int t) {
    switch (t) {
        case PNCU_INT8: case PNCU_UINT8: return launch_unary<F<uint8_t>>;
        case PNCU_INT16: case PNCU_UINT16: return launch_unary<F<uint16_t>>;
        case PNCU_INT32: case PNCU_UINT32: return launch_unary<F<uint32_t>>;
        case PNCU_INT64: case PNCU_UINT64: return launch_unary<F<uint64_t>>;
    }
    return NULL;
}
inline bool is_num(int t) { return (t >= PNCU_BOOL && t <= PNCU_UINT64) || t == PNCU_FLOAT || t == PNCU_DOUBLE || t == PNCU_LONGDOUBLE; }

}  // namespace

extern "C" pncu_unary_fn pncu_GetUnaryOpFast(int func, int t, int* wantedOutType) {
    if (!wantedOutType) return NULL;
    switch (func) {
        case PNCU_ABS:  // ops_unary.cpp:1636-1644, 1266-1278: every type hooked
            *wantedOutType = t;
            switch (t) {
                case PNCU_BOOL: case PNCU_UINT8: case PNCU_UINT16: case PNCU_UINT32: case PNCU_UINT64:
                    return pick_uint_width<IdentF>(t == PNCU_BOOL ? PNCU_UINT8 : t);
            }
            return pick_signed_fp<AbsF>(t);
        case PNCU_SIGNBIT:  // ops_unary.cpp:1612-1619
            if (t == PNCU_FLOAT || t == PNCU_DOUBLE || t == PNCU_LONGDOUBLE) {
                *wantedOutType = PNCU_BOOL;
                return pick_fp<SignBitF>(t);
            }
            return NULL;
        case PNCU_INVERT:
        case PNCU_BITWISE_NOT:  // ops_unary.cpp:1709-1715, 1328-1345
            if (t > PNCU_UINT64) return NULL;
            *wantedOutType = t;
            if (t == PNCU_BOOL) return launch_unary<BoolNotF>;
            return pick_uint_width<InvertF>(t);
        case PNCU_FLOOR:  // ops_unary.cpp:1716-1725
            if (t < PNCU_FLOAT) return NULL;
            *wantedOutType = t;
            return pick_fp<FloorF>(t);
        case PNCU_CEIL:
            if (t < PNCU_FLOAT) return NULL;
            *wantedOutType = t;
            return pick_fp<CeilF>(t);
        case PNCU_TRUNC:
            if (t < PNCU_FLOAT) return NULL;
            *wantedOutType = t;
            return pick_fp<TruncF>(t);
        case PNCU_ROUND:  // np.rint (src/pnumpy/_pnumpy.cpp:126); ops_unary.cpp:1746-1756
            if (t < PNCU_FLOAT) return NULL;
            *wantedOutType = t;
            return pick_fp<RintF>(t);
        case PNCU_NEGATIVE:  // ops_unary.cpp:1695-1708 (not on bool)
            if (t <= PNCU_BOOL) return NULL;
            *wantedOutType = t;
            if (t == PNCU_FLOAT || t == PNCU_DOUBLE) return pick_fp<NegF>(t);
            return pick_uint_width<NegF>(t);
        case PNCU_SIGN:  // ops_unary.cpp:1280-1292 (signed ints and floats)
            switch (t) {
                case PNCU_INT8: case PNCU_INT16: case PNCU_INT32: case PNCU_INT64:
                case PNCU_FLOAT: case PNCU_DOUBLE: case PNCU_LONGDOUBLE:
                    *wantedOutType = t;
                    return pick_signed_fp<SignF>(t);
            }
            return NULL;
        case PNCU_SQRT:  // ops_unary.cpp:1757-1770
            if (t < PNCU_FLOAT) return NULL;
            *wantedOutType = (t == PNCU_FLOAT) ? PNCU_FLOAT : PNCU_DOUBLE;
            return pick_fp<SqrtF>(t);
        case PNCU_LOGICAL_NOT:  // ops_unary.cpp:1308-1326
            *wantedOutType = PNCU_BOOL;
            return pick_all<LogicalNotF>(t);
        // classifiers: ops_unary.cpp:1646-1694 (fast) and 1347-1562 (slow); always -> bool
        case PNCU_ISNAN: *wantedOutType = PNCU_BOOL; return pick_classifier<IsNanF, 0>(t);
        case PNCU_ISNOTNAN: *wantedOutType = PNCU_BOOL; return pick_classifier<IsNotNanF, 1>(t);
        case PNCU_ISINF: *wantedOutType = PNCU_BOOL; return pick_classifier<IsInfF, 0>(t);
        case PNCU_ISNOTINF: *wantedOutType = PNCU_BOOL; return pick_classifier<IsNotInfF, 1>(t);
        case PNCU_ISFINITE: *wantedOutType = PNCU_BOOL; return pick_classifier<IsFiniteF, 1>(t);
        case PNCU_ISNOTFINITE: *wantedOutType = PNCU_BOOL; return pick_classifier<IsNotFiniteF, 0>(t);
        case PNCU_ISNORMAL: *wantedOutType = PNCU_BOOL; return pick_all<IsNormalF>(t);
        case PNCU_ISNOTNORMAL: *wantedOutType = PNCU_BOOL; return pick_all<IsNotNormalF>(t);
        case PNCU_ISNANORZERO: *wantedOutType = PNCU_BOOL; return pick_all<IsNanOrZeroF>(t);
    }
    (void)is_num;
    return NULL;
}
