// dtype conversion (`astype`): pncu_GetConvertFast.
//
// replaces: nothing the reference has yet -- conversion is the first item of its roadmap
// (doc_src/source/roadmap.rst:11-15); NumPy's own cast loops are what
// `ndarray.astype` runs today, single-threaded.  Same streaming kernels as the unary ops
// (ew1_vec_kernel with different In / Out widths: a lane moves 32 bytes of the wider type and
// 32 * narrow / wide bytes of the other).
//
// Semantics are NumPy's, i.e. C casts as x86-64 executes them (pinned by the oracle, which is gcc's
// code for the same casts, and by NumPy itself in the tests):
//   int -> int        wraps (truncation / sign extension);  anything -> bool is `x != 0` (NaN -> True)
//   int -> float      round to nearest even;  float64 -> float32 round to nearest even
//   float -> int      truncation toward zero.  Out of range / NaN is undefined in C; NumPy warns and
//                     returns what cvttss2si / cvttsd2si return, which is reproduced here:
//                       -> int32           0x80000000 ("integer indefinite")
//                       -> int64           0x8000000000000000
//                       -> int8/16, uint8/16   converted to int32 first, then truncated
//                       -> uint32          x < 2^31 (or NaN): as int32;  else (int32)(x - 2^31) ^ 2^31
//                       -> uint64          x < 2^63 (or NaN): as int64;  else (int64)(x - 2^63) ^ 2^63
#include "ew.cuh"
#include "functors.cuh"

using namespace pncu;

namespace {

// ---- float -> integer the way cvttss2si / cvttsd2si do it ------------------------------------------
template <class F> __device__ __forceinline__ int32_t x86_cvtt32(F x) {
    // in range iff -2^31 - 1 < x < 2^31 (truncation); NaN compares false
    if (x > (F)-2147483649.0 && x < (F)2147483648.0) return (int32_t)x;   // float: -2147483649.0f rounds to -2^31: see below
    return INT32_MIN;
}
template <> __device__ __forceinline__ int32_t x86_cvtt32<float>(float x) {
    // float32 cannot represent -2^31 - 1: the representable values below -2^31 are all out of range, -2^31 is in
    if (x >= -2147483648.0f && x < 2147483648.0f) return (int32_t)x;
    return INT32_MIN;
}
template <class F> __device__ __forceinline__ int64_t x86_cvtt64(F x) {
    // -2^63 is representable in both float types; everything below it is out of range
    if (x >= (F)-9223372036854775808.0 && x < (F)9223372036854775808.0) return (int64_t)x;
    return INT64_MIN;
}
template <class F> __device__ __forceinline__ uint32_t x86_cvttu32(F x) {
    // what NumPy's SSE-vectorised float -> uint32 loop does: 32-bit conversions on both sides of 2^31
    if (x >= (F)2147483648.0) return (uint32_t)x86_cvtt32<F>(x - (F)2147483648.0) ^ 0x80000000u;
    return (uint32_t)x86_cvtt32<F>(x);   // x < 2^31, or NaN
}
template <class F> __device__ __forceinline__ uint64_t x86_cvttu64(F x) {
    if (x >= (F)9223372036854775808.0) return (uint64_t)x86_cvtt64<F>(x - (F)9223372036854775808.0) ^ 0x8000000000000000ull;
    return (uint64_t)x86_cvtt64<F>(x);   // x < 2^63, or NaN
}

// IN_BOOL / OUT_BOOL: NumPy's bool is a uint8_t here (bool8), so the bool-ness of a side is a separate flag
template <class In_, class Out_, bool IN_BOOL, bool OUT_BOOL>
struct ConvF {
    typedef In_ In;
    typedef Out_ Out;
    static __device__ __forceinline__ Out apply(In a) {
        if constexpr (OUT_BOOL) {
            return (Out)(a != (In)0);
        } else if constexpr (IN_BOOL) {
            return (Out)(a != 0 ? 1 : 0);
        } else if constexpr (is_fp<In>::value && !is_fp<Out>::value) {
            if constexpr (std::is_same<Out, int64_t>::value) return x86_cvtt64<In>(a);
            else if constexpr (std::is_same<Out, uint64_t>::value) return x86_cvttu64<In>(a);
            else if constexpr (std::is_same<Out, uint32_t>::value) return x86_cvttu32<In>(a);
            else return (Out)(typename std::make_unsigned<Out>::type)(uint32_t)x86_cvtt32<In>(a);
        } else if constexpr (is_fp<In>::value && is_fp<Out>::value) {
            if constexpr (std::is_same<In, double>::value && std::is_same<Out, float>::value) return __double2float_rn(a);
            else return (Out)a;
        } else if constexpr (is_fp<Out>::value) {
            // int -> float: cvt.rn
            if constexpr (std::is_same<Out, float>::value) {
                if constexpr (std::is_signed<In>::value) return __ll2float_rn((long long)a);
                else return __ull2float_rn((unsigned long long)a);
            } else {
                if constexpr (std::is_signed<In>::value) return __ll2double_rn((long long)a);
                else return __ull2double_rn((unsigned long long)a);
            }
        } else {
            // int -> int: two's complement wrap
            typedef typename std::make_unsigned<Out>::type UO;
            return (Out)(UO)(unsigned long long)(long long)a;
        }
    }
};

template <class In, bool IN_BOOL> pncu_unary_fn pick_out(int out_t) {
    switch (out_t) {
        case PNCU_BOOL: return launch_unary<ConvF<In, bool8, IN_BOOL, true>>;
        case PNCU_INT8: return launch_unary<ConvF<In, int8_t, IN_BOOL, false>>;
        case PNCU_UINT8: return launch_unary<ConvF<In, uint8_t, IN_BOOL, false>>;
        case PNCU_INT16: return launch_unary<ConvF<In, int16_t, IN_BOOL, false>>;
        case PNCU_UINT16: return launch_unary<ConvF<In, uint16_t, IN_BOOL, false>>;
        case PNCU_INT32: return launch_unary<ConvF<In, int32_t, IN_BOOL, false>>;
        case PNCU_UINT32: return launch_unary<ConvF<In, uint32_t, IN_BOOL, false>>;
        case PNCU_INT64: return launch_unary<ConvF<In, int64_t, IN_BOOL, false>>;
        case PNCU_UINT64: return launch_unary<ConvF<In, uint64_t, IN_BOOL, false>>;
        case PNCU_FLOAT: return launch_unary<ConvF<In, float, IN_BOOL, false>>;
        case PNCU_DOUBLE: return launch_unary<ConvF<In, double, IN_BOOL, false>>;
    }
    return NULL;
}

}  // namespace

// out[i] = (OutType)in[i] with NumPy's cast semantics; NULL when either type has no kernel
extern "C" pncu_unary_fn pncu_GetConvertFast(int in_t, int out_t) {
    switch (in_t) {
        case PNCU_BOOL: return pick_out<bool8, true>(out_t);
        case PNCU_INT8: return pick_out<int8_t, false>(out_t);
        case PNCU_UINT8: return pick_out<uint8_t, false>(out_t);
        case PNCU_INT16: return pick_out<int16_t, false>(out_t);
        case PNCU_UINT16: return pick_out<uint16_t, false>(out_t);
        case PNCU_INT32: return pick_out<int32_t, false>(out_t);
        case PNCU_UINT32: return pick_out<uint32_t, false>(out_t);
        case PNCU_INT64: return pick_out<int64_t, false>(out_t);
        case PNCU_UINT64: return pick_out<uint64_t, false>(out_t);
        case PNCU_FLOAT: return pick_out<float, false>(out_t);
        case PNCU_DOUBLE: return pick_out<double, false>(out_t);
    }
    return NULL;
}
