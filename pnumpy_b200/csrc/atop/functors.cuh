// Per-element operations.  Each functor states the IEEE / integer semantics it must
// reproduce and the reference loop body it replaces.  Device code is compiled with
// -fmad=false; float add/sub/mul/div/sqrt additionally go through the *_rn intrinsics,
// which ptxas never contracts or approximates, so results are the correctly rounded
// IEEE-754 values that the reference's vaddps/vmulps/vdivps/vsqrtps produce.
#pragma once
#include <stdint.h>
#include <type_traits>

namespace pncu {

typedef uint8_t bool8;  // NumPy bool: one byte, canonical values 0/1

template <class T> struct is_fp { static constexpr bool value = std::is_floating_point<T>::value; };

// ---------------------------------------------------------------------------------------
// arithmetic.  Integers wrap modulo 2^k (computed unsigned to stay defined).
// reference: AddOp/SubOp/MulOp/DivOp + ADD_OP_256* etc, src/atop/ops_binary.cpp:40-185
// ---------------------------------------------------------------------------------------
template <class T> __device__ __forceinline__ T op_add(T a, T b) {
    typedef typename std::make_unsigned<T>::type U;
    return (T)(U)((U)a + (U)b);
}
template <> __device__ __forceinline__ float op_add<float>(float a, float b) { return __fadd_rn(a, b); }
template <> __device__ __forceinline__ double op_add<double>(double a, double b) { return __dadd_rn(a, b); }

template <class T> __device__ __forceinline__ T op_sub(T a, T b) {
    typedef typename std::make_unsigned<T>::type U;
    return (T)(U)((U)a - (U)b);
}
template <> __device__ __forceinline__ float op_sub<float>(float a, float b) { return __fsub_rn(a, b); }
template <> __device__ __forceinline__ double op_sub<double>(double a, double b) { return __dsub_rn(a, b); }

template <class T> __device__ __forceinline__ T op_mul(T a, T b) {
    typedef typename std::make_unsigned<T>::type U;
    // widen sub-int types explicitly: uint16*uint16 promotes to (signed) int and may overflow
    typedef typename std::conditional<(sizeof(T) < 4), unsigned int, U>::type W;
    return (T)(U)((W)(U)a * (W)(U)b);
}
template <> __device__ __forceinline__ float op_mul<float>(float a, float b) { return __fmul_rn(a, b); }
template <> __device__ __forceinline__ double op_mul<double>(double a, double b) { return __dmul_rn(a, b); }

__device__ __forceinline__ float op_div(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double op_div(double a, double b) { return __ddiv_rn(a, b); }

// NumPy floor_divide for floats (npy_floor_divide -> npy_divmod, numpy/_core/src/npymath/
// npy_math_internal.h.src): quotient from an exact fmod, Python sign convention, snapped to the
// nearest integer.  The reference's scalar body is floor(x / y) (src/atop/ops_binary.cpp:80-82),
// which differs from NumPy by one when x / y rounds up to an integer; NumPy semantics are used.
__device__ __forceinline__ float fp_fmod(float a, float b) { return fmodf(a, b); }
__device__ __forceinline__ double fp_fmod(double a, double b) { return fmod(a, b); }
__device__ __forceinline__ float fp_floor(float a) { return floorf(a); }
__device__ __forceinline__ double fp_floor(double a) { return floor(a); }
__device__ __forceinline__ float fp_copysign(float a, float b) { return copysignf(a, b); }
__device__ __forceinline__ double fp_copysign(double a, double b) { return copysign(a, b); }
template <class T> __device__ __forceinline__ T op_floordiv(T a, T b) {
    if (b == (T)0) return op_div(a, b);
    const T mod = fp_fmod(a, b);
    T div = op_div(op_sub<T>(a, mod), b);
    if (mod != (T)0) {
        if ((b < (T)0) != (mod < (T)0)) div = op_sub<T>(div, (T)1);
    }
    if (div != (T)0) {
        T fl = fp_floor(div);
        if (op_sub<T>(div, fl) > (T)0.5) fl = op_add<T>(fl, (T)1);
        return fl;
    }
    return fp_copysign((T)0, op_div(a, b));
}

// NumPy minimum/maximum: NaN-propagating, and for equal operands (incl. +0/-0) the
// SECOND operand is returned -- `(a < b || isnan(a)) ? a : b`.  The reference's float
// body is _mm256_min_ps (src/atop/ops_binary.cpp:146-147) which drops a NaN in the
// first operand; SURVEY.md 0.5 marks that a defect, so NumPy semantics are used.
template <class T> __device__ __forceinline__ T op_min(T a, T b) {
    if constexpr (is_fp<T>::value) return (a < b || a != a) ? a : b;
    else return a < b ? a : b;
}
template <class T> __device__ __forceinline__ T op_max(T a, T b) {
    if constexpr (is_fp<T>::value) return (a > b || a != a) ? a : b;
    else return a > b ? a : b;
}

// NumPy shifts (npy_lshift / npy_rshift): a count outside [0, bits) gives 0, or -1 for
// an arithmetic right shift of a negative value.
template <class T> __device__ __forceinline__ T op_lshift(T a, T b) {
    typedef typename std::make_unsigned<T>::type U;
    return ((U)b < (U)(sizeof(T) * 8)) ? (T)(U)((U)a << (U)b) : (T)0;
}
template <class T> __device__ __forceinline__ T op_rshift(T a, T b) {
    typedef typename std::make_unsigned<T>::type U;
    if ((U)b < (U)(sizeof(T) * 8)) return (T)(a >> b);
    if constexpr (std::is_signed<T>::value) return a < 0 ? (T)-1 : (T)0;
    else return (T)0;
}

#define PNCU_BINARY_FUNCTOR(NAME, EXPR)                                        \
    template <class T> struct NAME {                                           \
        typedef T In;                                                          \
        typedef T Out;                                                         \
        static __device__ __forceinline__ T apply(T a, T b) { return EXPR; }   \
    };

PNCU_BINARY_FUNCTOR(AddF, op_add<T>(a, b))
PNCU_BINARY_FUNCTOR(SubF, op_sub<T>(a, b))
PNCU_BINARY_FUNCTOR(MulF, op_mul<T>(a, b))
PNCU_BINARY_FUNCTOR(DivF, op_div(a, b))
PNCU_BINARY_FUNCTOR(FloorDivF, op_floordiv<T>(a, b))
PNCU_BINARY_FUNCTOR(MinF, op_min<T>(a, b))
PNCU_BINARY_FUNCTOR(MaxF, op_max<T>(a, b))
PNCU_BINARY_FUNCTOR(AndF, (T)(a & b))
PNCU_BINARY_FUNCTOR(OrF, (T)(a | b))
PNCU_BINARY_FUNCTOR(XorF, (T)(a ^ b))
PNCU_BINARY_FUNCTOR(LShiftF, op_lshift<T>(a, b))
PNCU_BINARY_FUNCTOR(RShiftF, op_rshift<T>(a, b))

// bool: NumPy's BOOL_add/maximum/logical_or/bitwise_or are `in1 || in2`, BOOL_multiply/
// minimum/logical_and/bitwise_and are `in1 && in2`, always yielding 0/1
// (the reference ORs/ANDs the raw bytes: src/atop/ops_binary.cpp:878,892)
struct BoolOrF  { typedef bool8 In; typedef bool8 Out; static __device__ __forceinline__ bool8 apply(bool8 a, bool8 b) { return (bool8)((a != 0) | (b != 0)); } };
struct BoolAndF { typedef bool8 In; typedef bool8 Out; static __device__ __forceinline__ bool8 apply(bool8 a, bool8 b) { return (bool8)((a != 0) & (b != 0)); } };
struct BoolXorF { typedef bool8 In; typedef bool8 Out; static __device__ __forceinline__ bool8 apply(bool8 a, bool8 b) { return (bool8)((a != 0) ^ (b != 0)); } };

// int / int -> float64 true divide: cvt.rn.f64 of both operands, then div.rn.f64
// (what NumPy does after casting both operands to double).
template <class T> struct IntTrueDivF {
    typedef T In;
    typedef double Out;
    static __device__ __forceinline__ double apply(T a, T b) { return __ddiv_rn((double)a, (double)b); }
};

// ---------------------------------------------------------------------------------------
// comparisons -> bool.  Float predicates are the ordered ones, != is unordered, exactly
// what C's operators give (reference: _CMP_*_OS / _CMP_NEQ_US, src/atop/ops_compare.cpp:
// 1303-1318).  Bool inputs are clamped to {0,1} first (src/atop/ops_compare.cpp:317-323).
// ---------------------------------------------------------------------------------------
template <class T, int OP> struct CmpF {
    typedef T In;
    typedef bool8 Out;
    static __device__ __forceinline__ bool8 apply(T a, T b) {
        switch (OP) {
            case PNCU_CMP_EQ: return a == b;
            case PNCU_CMP_NE: return a != b;
            case PNCU_CMP_LT: return a < b;
            case PNCU_CMP_GT: return a > b;
            case PNCU_CMP_LTE: return a <= b;
            default: return a >= b;
        }
    }
};
template <int OP> struct CmpBoolF {
    typedef bool8 In;
    typedef bool8 Out;
    static __device__ __forceinline__ bool8 apply(bool8 a, bool8 b) {
        return CmpF<int, OP>::apply((int)(a != 0), (int)(b != 0));
    }
};

// ---------------------------------------------------------------------------------------
// Word-at-a-time forms for the 1- and 2-byte types.  Taking a 32-byte pack apart element by element
// costs 3-4 ALU instructions per byte, which makes the int8/uint8/bool kernels instruction bound
// (int8 add: 78 % of the HBM rate where int32 add reaches 86 %).  The SIMD-in-a-word intrinsics
// (__vadd4, __vsetgts4, ... : per-lane wrap-around arithmetic and 0/1 predicates) do 4 or 2 elements
// per instruction group and give the same bits lane for lane, so ew.cuh applies WordOp<F>::word to
// whole 32-bit words when it exists.  The per-element apply() stays the definition (and the body of
// the scalar head/tail and the strided kernels).
// ---------------------------------------------------------------------------------------
template <class F> struct WordOp { static constexpr bool enabled = false; };
template <class F> struct WordOp1 { static constexpr bool enabled = false; };   // unary

#define PNCU_WORD2(FUNCTOR, EXPR)                                                                   \
    template <> struct WordOp<FUNCTOR> {                                                            \
        static constexpr bool enabled = true;                                                       \
        static __device__ __forceinline__ unsigned word(unsigned a, unsigned b) { return EXPR; }    \
    };
// wrap-around add / subtract, bitwise: the same for signed and unsigned lanes
#define PNCU_WORD2_BOTH(NAME, E4, E2)                                                               \
    PNCU_WORD2(NAME<int8_t>, E4) PNCU_WORD2(NAME<uint8_t>, E4) PNCU_WORD2(NAME<int16_t>, E2) PNCU_WORD2(NAME<uint16_t>, E2)
PNCU_WORD2_BOTH(AddF, __vadd4(a, b), __vadd2(a, b))
PNCU_WORD2_BOTH(SubF, __vsub4(a, b), __vsub2(a, b))
PNCU_WORD2_BOTH(AndF, a & b, a & b)
PNCU_WORD2_BOTH(OrF, a | b, a | b)
PNCU_WORD2_BOTH(XorF, a ^ b, a ^ b)
PNCU_WORD2(MinF<int8_t>, __vmins4(a, b)) PNCU_WORD2(MinF<uint8_t>, __vminu4(a, b))
PNCU_WORD2(MinF<int16_t>, __vmins2(a, b)) PNCU_WORD2(MinF<uint16_t>, __vminu2(a, b))
PNCU_WORD2(MaxF<int8_t>, __vmaxs4(a, b)) PNCU_WORD2(MaxF<uint8_t>, __vmaxu4(a, b))
PNCU_WORD2(MaxF<int16_t>, __vmaxs2(a, b)) PNCU_WORD2(MaxF<uint16_t>, __vmaxu2(a, b))
// bool: canonical 0/1 out of any non-zero byte
PNCU_WORD2(BoolOrF, __vsetne4(a | b, 0u))
PNCU_WORD2(BoolAndF, __vsetne4(a, 0u) & __vsetne4(b, 0u))
PNCU_WORD2(BoolXorF, __vsetne4(a, 0u) ^ __vsetne4(b, 0u))
// byte comparisons -> bool bytes (1-byte inputs only: input and output packs have the same shape)
#define PNCU_WORD_CMP(OP, S4, U4)                                                                   \
    PNCU_WORD2(CmpF<int8_t PNCU_COMMA OP>, S4(a, b)) PNCU_WORD2(CmpF<uint8_t PNCU_COMMA OP>, U4(a, b)) \
    PNCU_WORD2(CmpBoolF<OP>, U4(__vsetne4(a, 0u), __vsetne4(b, 0u)))
#define PNCU_COMMA ,
PNCU_WORD_CMP(PNCU_CMP_EQ, __vseteq4, __vseteq4)
PNCU_WORD_CMP(PNCU_CMP_NE, __vsetne4, __vsetne4)
PNCU_WORD_CMP(PNCU_CMP_LT, __vsetlts4, __vsetltu4)
PNCU_WORD_CMP(PNCU_CMP_GT, __vsetgts4, __vsetgtu4)
PNCU_WORD_CMP(PNCU_CMP_LTE, __vsetles4, __vsetleu4)
PNCU_WORD_CMP(PNCU_CMP_GTE, __vsetges4, __vsetgeu4)
#undef PNCU_WORD_CMP
#undef PNCU_WORD2_BOTH
#undef PNCU_WORD2

// a scalar operand replicated into every lane of a word
template <class T> __device__ __forceinline__ unsigned splat_word(T v) {
    if constexpr (sizeof(T) == 1) return (unsigned)(uint8_t)v * 0x01010101u;
    else if constexpr (sizeof(T) == 2) return (unsigned)(uint16_t)v * 0x00010001u;
    else return 0u;
}

}  // namespace pncu
