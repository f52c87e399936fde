/*
 * atop_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain scalar C restatement of pnumpy's atop ufunc hot path (the loop bodies behind
 * src/atop/atop.h:19-29) used ONLY as the checker for the CUDA kernels: by tests/, by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs.  Nothing
 * under pnumpy_b200/ may import, link or call it; the product has no CPU fallback.
 *
 * Pinning.  tests/test_oracle.py checks every function here against
 *   (a) the reference's own AVX2 loop bodies, compiled unmodified from /root/reference/src/atop
 *       into oracle/_ref/libatop_ref.so by oracle/Makefile (bit-exact wherever the reference is
 *       not defective), and
 *   (b) golden vectors under tests/golden/ generated from that library and from NumPy 2.3.5 by
 *       tests/golden/make_golden.py (these travel to the GPU box; /root/reference does not).
 * Where the reference is numerically defective the oracle follows
 * NumPy / IEEE-754 and the header of each function says so.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math (see oracle/Makefile) so that no FMA is
 * formed and every float operation is a single correctly rounded IEEE operation.
 *
 * All entry points take the reference's argument conventions (byte strides, stride 0 =
 * broadcast, enum values of src/atop/common_inc.h:236-372) and return 0, or -1 when the
 * (op, type) pair is outside the hot path.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

enum { T_BOOL = 0, T_I8, T_U8, T_I16, T_U16, T_I32, T_U32, T_I64, T_U64, T_F32 = 12, T_F64 = 13 };
enum { B_ADD = 1, B_SUB = 2, B_MUL = 3, B_MIN = 5, B_MAX = 6, B_FLOORDIV = 9, B_DIV = 13, B_LAND = 16, B_LOR = 18,
       B_LSHIFT = 19, B_RSHIFT = 20, B_AND = 21, B_XOR = 22, B_OR = 23 };
enum { C_EQ = 0, C_NE, C_LT, C_GT, C_LE, C_GE };
enum { U_ABS = 1, U_SIGNBIT = 2, U_INVERT = 4, U_FLOOR = 5, U_CEIL = 6, U_TRUNC = 7, U_ROUND = 8,
       U_NEG = 9, U_SIGN = 11, U_SQRT = 15, U_LNOT = 18, U_ISINF = 19, U_ISNAN = 20, U_ISFINITE = 21,
       U_BITNOT = 28 };
enum { G_SIN = 1, G_COS = 2, G_TAN = 3, G_ASIN = 4, G_ACOS = 5, G_ATAN = 6, G_SINH = 7, G_COSH = 8, G_TANH = 9,
       G_ASINH = 10, G_ACOSH = 11, G_ATANH = 12, G_LOG = 13, G_LOG2 = 14, G_LOG10 = 15, G_EXP = 16, G_EXP2 = 17,
       G_EXPM1 = 18, G_LOG1P = 19, G_CBRT = 20 };

int oracle_itemsize(int t) {
    switch (t) {
        case T_BOOL: case T_I8: case T_U8: return 1;
        case T_I16: case T_U16: return 2;
        case T_I32: case T_U32: case T_F32: return 4;
        case T_I64: case T_U64: case T_F64: return 8;
    }
    return 0;
}

#define AT(T, p, i, s) (*(const T*)((const char*)(p) + (i) * (s)))
#define OUT(T, p, i, s) (*(T*)((char*)(p) + (i) * (s)))

/* ------------------------------------------------------------------------------------------
 * binary: out[i] = in1[i] op in2[i]
 * follows SimpleMathOpFast / ...Symmetric / ...Slow, src/atop/ops_binary.cpp:349-793 and the
 * op functors :40-120.  Integers wrap; float ops are single IEEE operations.
 * Deviation: float MIN/MAX propagate NaN like NumPy; bool ADD/MUL/AND/OR/XOR
 * yield canonical 0/1 like NumPy (the reference ORs/ANDs raw bytes, same on canonical input).
 * ---------------------------------------------------------------------------------------- */
#define BIN_LOOP(T, EXPR)                                                      \
    for (int64_t i = 0; i < n; ++i) {                                          \
        const T a = AT(T, in1, i, s1), b = AT(T, in2, i, s2);                  \
        OUT(T, out, i, so) = (T)(EXPR);                                        \
    }                                                                          \
    return 0;

#define INT_BIN(T, U, BITS)                                                    \
    switch (op) {                                                              \
        case B_ADD: BIN_LOOP(T, (U)((U)a + (U)b))                              \
        case B_SUB: BIN_LOOP(T, (U)((U)a - (U)b))                              \
        case B_MUL: BIN_LOOP(T, (U)((uint64_t)(U)a * (uint64_t)(U)b))          \
        case B_MIN: BIN_LOOP(T, a < b ? a : b)                                 \
        case B_MAX: BIN_LOOP(T, a > b ? a : b)                                 \
        case B_AND: BIN_LOOP(T, a & b)                                         \
        case B_OR: BIN_LOOP(T, a | b)                                          \
        case B_XOR: BIN_LOOP(T, a ^ b)                                         \
        case B_LSHIFT: BIN_LOOP(T, ((U)b < BITS) ? (U)((U)a << (U)b) : 0)      \
        case B_RSHIFT: BIN_LOOP(T, ((U)b < BITS) ? (a >> b) : (a < 0 ? -1 : 0)) \
    }                                                                          \
    return -1;

/* NumPy's npy_floor_divide / npy_divmod (numpy/_core/src/npymath/npy_math_internal.h.src).
 * The reference's scalar body is floor(x / y) (src/atop/ops_binary.cpp:80-82); they differ by
 * one where x / y rounds up to an integer, and NumPy's value is the drop-in contract. */
#define FLOORDIV_FN(NAME, T, SFX)                                              \
    static T NAME(T a, T b) {                                                  \
        if (b == 0) return a / b;                                              \
        const T mod = fmod##SFX(a, b);                                         \
        T div = (a - mod) / b;                                                 \
        if (mod != 0 && ((b < 0) != (mod < 0))) div -= 1;                      \
        if (div != 0) {                                                        \
            T fl = floor##SFX(div);                                            \
            if (div - fl > (T)0.5) fl += 1;                                    \
            return fl;                                                         \
        }                                                                      \
        return copysign##SFX(0, a / b);                                        \
    }
FLOORDIV_FN(floordiv_f32, float, f)
FLOORDIV_FN(floordiv_f64, double, )
#define FLOORDIV(a, b) _Generic((a), float: floordiv_f32, double: floordiv_f64)(a, b)

#define FP_BIN(T)                                                              \
    switch (op) {                                                              \
        case B_FLOORDIV: BIN_LOOP(T, FLOORDIV(a, b))                           \
        case B_ADD: BIN_LOOP(T, a + b)                                         \
        case B_SUB: BIN_LOOP(T, a - b)                                         \
        case B_MUL: BIN_LOOP(T, a * b)                                         \
        case B_DIV: BIN_LOOP(T, a / b)                                         \
        case B_MIN: BIN_LOOP(T, (a < b || a != a) ? a : b)                     \
        case B_MAX: BIN_LOOP(T, (a > b || a != a) ? a : b)                     \
    }                                                                          \
    return -1;

int oracle_binary(int op, int t, const void* in1, const void* in2, void* out, int64_t n,
                  int64_t s1, int64_t s2, int64_t so) {
    switch (t) {
        case T_BOOL:
            switch (op) {
                case B_ADD: case B_MAX: case B_LOR: case B_OR: BIN_LOOP(uint8_t, (a != 0) | (b != 0))
                case B_MUL: case B_MIN: case B_LAND: case B_AND: BIN_LOOP(uint8_t, (a != 0) & (b != 0))
                case B_XOR: BIN_LOOP(uint8_t, (a != 0) ^ (b != 0))
            }
            return -1;
        case T_I8: INT_BIN(int8_t, uint8_t, 8)
        case T_U8: INT_BIN(uint8_t, uint8_t, 8)
        case T_I16: INT_BIN(int16_t, uint16_t, 16)
        case T_U16: INT_BIN(uint16_t, uint16_t, 16)
        case T_I32: INT_BIN(int32_t, uint32_t, 32)
        case T_U32: INT_BIN(uint32_t, uint32_t, 32)
        case T_I64: INT_BIN(int64_t, uint64_t, 64)
        case T_U64: INT_BIN(uint64_t, uint64_t, 64)
        case T_F32: FP_BIN(float)
        case T_F64: FP_BIN(double)
    }
    return -1;
}

/* int / int -> float64: what NumPy's dd->d loop sees after casting */
#define IDIV_LOOP(T)                                                           \
    for (int64_t i = 0; i < n; ++i)                                            \
        OUT(double, out, i, so) = (double)AT(T, in1, i, s1) / (double)AT(T, in2, i, s2); \
    return 0;
int oracle_int_true_divide(int t, const void* in1, const void* in2, void* out, int64_t n,
                           int64_t s1, int64_t s2, int64_t so) {
    switch (t) {
        case T_I8: IDIV_LOOP(int8_t)
        case T_U8: IDIV_LOOP(uint8_t)
        case T_I16: IDIV_LOOP(int16_t)
        case T_U16: IDIV_LOOP(uint16_t)
        case T_I32: IDIV_LOOP(int32_t)
        case T_U32: IDIV_LOOP(uint32_t)
        case T_I64: IDIV_LOOP(int64_t)
        case T_U64: IDIV_LOOP(uint64_t)
    }
    return -1;
}

/* ------------------------------------------------------------------------------------------
 * compare: out[i] = (in1[i] cmp in2[i]) as one bool byte
 * follows Compare* and GetComparisonOpFast, src/atop/ops_compare.cpp:400-1446: ordered
 * predicates, unordered !=, bool inputs clamped to {0,1} (:317-323).
 * ---------------------------------------------------------------------------------------- */
#define CMP_LOOP(T, EXPR)                                                      \
    for (int64_t i = 0; i < n; ++i) {                                          \
        const T a = AT(T, in1, i, s1), b = AT(T, in2, i, s2);                  \
        OUT(uint8_t, out, i, so) = (uint8_t)(EXPR);                            \
    }                                                                          \
    return 0;
#define CMP_TYPE(T)                                                            \
    switch (op) {                                                              \
        case C_EQ: CMP_LOOP(T, a == b)                                         \
        case C_NE: CMP_LOOP(T, a != b)                                         \
        case C_LT: CMP_LOOP(T, a < b)                                          \
        case C_GT: CMP_LOOP(T, a > b)                                          \
        case C_LE: CMP_LOOP(T, a <= b)                                         \
        case C_GE: CMP_LOOP(T, a >= b)                                         \
    }                                                                          \
    return -1;
int oracle_compare(int op, int t, const void* in1, const void* in2, void* out, int64_t n,
                   int64_t s1, int64_t s2, int64_t so) {
    switch (t) {
        case T_BOOL:
            switch (op) {
                case C_EQ: CMP_LOOP(uint8_t, (a != 0) == (b != 0))
                case C_NE: CMP_LOOP(uint8_t, (a != 0) != (b != 0))
                case C_LT: CMP_LOOP(uint8_t, (a != 0) < (b != 0))
                case C_GT: CMP_LOOP(uint8_t, (a != 0) > (b != 0))
                case C_LE: CMP_LOOP(uint8_t, (a != 0) <= (b != 0))
                case C_GE: CMP_LOOP(uint8_t, (a != 0) >= (b != 0))
            }
            return -1;
        case T_I8: CMP_TYPE(int8_t)
        case T_U8: CMP_TYPE(uint8_t)
        case T_I16: CMP_TYPE(int16_t)
        case T_U16: CMP_TYPE(uint16_t)
        case T_I32: CMP_TYPE(int32_t)
        case T_U32: CMP_TYPE(uint32_t)
        case T_I64: CMP_TYPE(int64_t)
        case T_U64: CMP_TYPE(uint64_t)
        case T_F32: CMP_TYPE(float)
        case T_F64: CMP_TYPE(double)
    }
    return -1;
}

/* ------------------------------------------------------------------------------------------
 * unary
 * follows UnaryOpFast + the scalar fallbacks, src/atop/ops_unary.cpp:274-376, 435-1204.
 * Deviation: float NEGATIVE flips the sign bit like NumPy (the reference
 * computes 0 - x, :295-305).
 * ---------------------------------------------------------------------------------------- */
#define UN_LOOP(T, TO, EXPR)                                                   \
    for (int64_t i = 0; i < n; ++i) {                                          \
        const T a = AT(T, in, i, s1); (void)a;                                 \
        OUT(TO, out, i, so) = (TO)(EXPR);                                      \
    }                                                                          \
    return 0;
#define UN_SINT(T, U)                                                          \
    switch (op) {                                                              \
        case U_ABS: UN_LOOP(T, T, a < 0 ? (U)((U)0 - (U)a) : (U)a)             \
        case U_NEG: UN_LOOP(T, T, (U)((U)0 - (U)a))                            \
        case U_SIGN: UN_LOOP(T, T, a > 0 ? 1 : (a < 0 ? -1 : 0))               \
        case U_INVERT: case U_BITNOT: UN_LOOP(T, T, ~a)                        \
        case U_LNOT: UN_LOOP(T, uint8_t, a == 0)                               \
        case U_ISNAN: case U_ISINF: UN_LOOP(T, uint8_t, 0)                     \
        case U_ISFINITE: UN_LOOP(T, uint8_t, 1)                                \
    }                                                                          \
    return -1;
#define UN_UINT(T)                                                             \
    switch (op) {                                                              \
        case U_ABS: UN_LOOP(T, T, a)                                           \
        case U_NEG: UN_LOOP(T, T, (T)((T)0 - a))                               \
        case U_INVERT: case U_BITNOT: UN_LOOP(T, T, ~a)                        \
        case U_LNOT: UN_LOOP(T, uint8_t, a == 0)                               \
        case U_ISNAN: case U_ISINF: UN_LOOP(T, uint8_t, 0)                     \
        case U_ISFINITE: UN_LOOP(T, uint8_t, 1)                                \
    }                                                                          \
    return -1;
#define UN_FP(T, SFX)                                                          \
    switch (op) {                                                              \
        case U_ABS: UN_LOOP(T, T, fabs##SFX(a))                                \
        case U_NEG: UN_LOOP(T, T, -a)                                          \
        case U_SIGN: UN_LOOP(T, T, a > 0 ? 1 : (a < 0 ? -1 : (a == 0 ? 0 : a))) \
        case U_FLOOR: UN_LOOP(T, T, floor##SFX(a))                             \
        case U_CEIL: UN_LOOP(T, T, ceil##SFX(a))                               \
        case U_TRUNC: UN_LOOP(T, T, trunc##SFX(a))                             \
        case U_ROUND: UN_LOOP(T, T, rint##SFX(a))                              \
        case U_SQRT: UN_LOOP(T, T, sqrt##SFX(a))                               \
        case U_LNOT: UN_LOOP(T, uint8_t, a == 0)                               \
        case U_ISNAN: UN_LOOP(T, uint8_t, a != a)                              \
        case U_ISINF: UN_LOOP(T, uint8_t, fabs##SFX(a) == (T)INFINITY)         \
        case U_ISFINITE: UN_LOOP(T, uint8_t, fabs##SFX(a) < (T)INFINITY)       \
        case U_SIGNBIT: UN_LOOP(T, uint8_t, signbit(a) != 0)                   \
    }                                                                          \
    return -1;
int oracle_unary(int op, int t, const void* in, void* out, int64_t n, int64_t s1, int64_t so) {
    switch (t) {
        case T_BOOL:
            switch (op) {
                case U_ABS: UN_LOOP(uint8_t, uint8_t, a)
                case U_INVERT: case U_BITNOT: case U_LNOT: UN_LOOP(uint8_t, uint8_t, a == 0)
                case U_ISNAN: case U_ISINF: UN_LOOP(uint8_t, uint8_t, 0)
                case U_ISFINITE: UN_LOOP(uint8_t, uint8_t, 1)
            }
            return -1;
        case T_I8: UN_SINT(int8_t, uint8_t)
        case T_I16: UN_SINT(int16_t, uint16_t)
        case T_I32: UN_SINT(int32_t, uint32_t)
        case T_I64: UN_SINT(int64_t, uint64_t)
        case T_U8: UN_UINT(uint8_t)
        case T_U16: UN_UINT(uint16_t)
        case T_U32: UN_UINT(uint32_t)
        case T_U64: UN_UINT(uint64_t)
        case T_F32: UN_FP(float, f)
        case T_F64: UN_FP(double, )
    }
    return -1;
}

/* ------------------------------------------------------------------------------------------
 * float32 sin / cos / log / exp: the Cephes operation sequence of
 * src/atop/avx_mathfun.h:177-248 (log256_ps), :354-409 (exp256_ps), :506-601 (sin256_ps),
 * :604-673 (cos256_ps), one lane at a time.  No FMA, no division: bit-identical to the
 * reference's vector body on its valid domain.  Outside that domain the reference is
 * defective and this follows IEEE / NumPy: log(+-0) = -inf, log(<0) = nan,
 * log(inf) = inf, denormal inputs rescaled; exp overflow -> inf, underflow through rounded
 * denormals; |x| > 8192 or non-finite sin/cos through libm sinf/cosf.
 * (The reference also runs libm on the len%8 tail, src/atop/ops_trig.cpp:271-274; the oracle
 * is position independent.)
 * ---------------------------------------------------------------------------------------- */
static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

float oracle_logf(float x) {
    if (!(x > 0.0f)) return (x == 0.0f) ? -INFINITY : NAN;
    if (x == INFINITY) return INFINITY;
    int eadj = 0;
    if (x < FLT_MIN) { x = x * 8388608.0f; eadj = -23; }
    const uint32_t bits = f2u(x);
    const int imm0 = (int)(bits >> 23) - 0x7f + eadj;
    x = u2f((bits & 0x807fffffu) | 0x3f000000u);
    float e = (float)imm0 + 1.0f;
    const int lt = x < 0.707106781186547524f;
    const float tmp0 = lt ? x : 0.0f;
    x = x - 1.0f;
    e = e - (lt ? 1.0f : 0.0f);
    x = x + tmp0;
    const float z = x * x;
    float y = 7.0376836292E-2f;
    y = y * x; y = y + -1.1514610310E-1f;
    y = y * x; y = y + 1.1676998740E-1f;
    y = y * x; y = y + -1.2420140846E-1f;
    y = y * x; y = y + 1.4249322787E-1f;
    y = y * x; y = y + -1.6668057665E-1f;
    y = y * x; y = y + 2.0000714765E-1f;
    y = y * x; y = y + -2.4999993993E-1f;
    y = y * x; y = y + 3.3333331174E-1f;
    y = y * x;
    y = y * z;
    float tmp = e * -2.12194440e-4f;
    y = y + tmp;
    tmp = z * 0.5f;
    y = y - tmp;
    tmp = e * 0.693359375f;
    x = x + y;
    x = x + tmp;
    return x;
}

float oracle_expf(float x) {
    if (x != x) return x;
    if (x > 88.72283935546875f) return INFINITY;
    if (x < -103.98f) return 0.0f;
    float fx = x * 1.44269504088896341f;
    fx = fx + 0.5f;
    fx = floorf(fx);
    float tmp = fx * 0.693359375f;
    float z = fx * -2.12194440e-4f;
    x = x - tmp;
    x = x - z;
    z = x * x;
    float y = 1.9875691500E-4f;
    y = y * x; y = y + 1.3981999507E-3f;
    y = y * x; y = y + 8.3334519073E-3f;
    y = y * x; y = y + 4.1665795894E-2f;
    y = y * x; y = y + 1.6666665459E-1f;
    y = y * x; y = y + 5.0000001201E-1f;
    y = y * z;
    y = y + x;
    y = y + 1.0f;
    const int n = (int)fx;
    if (n > 127) { y = y * u2f((uint32_t)(n - 1 + 0x7f) << 23); return y * 2.0f; }
    if (n < -126) { y = y * u2f((uint32_t)(n + 24 + 0x7f) << 23); return y * 5.9604644775390625e-08f; }
    return y * u2f((uint32_t)(n + 0x7f) << 23);
}

static float sincos_tail(float x, float y, int sin_poly, uint32_t sign) {
    const float x1 = y * -0.78515625f;
    const float x2 = y * -2.4187564849853515625e-4f;
    const float x3 = y * -3.77489497744594108e-8f;
    x = x + x1;
    x = x + x2;
    x = x + x3;
    const float z = x * x;
    float c = 2.443315711809948E-005f;
    c = c * z; c = c + -1.388731625493765E-003f;
    c = c * z; c = c + 4.166664568298827E-002f;
    c = c * z;
    c = c * z;
    c = c - z * 0.5f;
    c = c + 1.0f;
    float s = -1.9515295891E-4f;
    s = s * z; s = s + 8.3321608736E-3f;
    s = s * z; s = s + -1.6666654611E-1f;
    s = s * z;
    s = s * x;
    s = s + x;
    const float r = sin_poly ? (0.0f + s) : (c + 0.0f);
    return u2f(f2u(r) ^ sign);
}

float oracle_sinf(float xin) {
    const float x = fabsf(xin);
    if (!(x <= 8192.0f)) return sinf(xin);
    uint32_t sign = f2u(xin) & 0x80000000u;
    float y = x * 1.27323954473516f;
    int j = (int)y;
    j = (j + 1) & ~1;
    y = (float)j;
    sign ^= ((uint32_t)(j & 4)) << 29;
    return sincos_tail(x, y, (j & 2) == 0, sign);
}

float oracle_cosf(float xin) {
    const float x = fabsf(xin);
    if (!(x <= 8192.0f)) return cosf(xin);
    float y = x * 1.27323954473516f;
    int j = (int)y;
    j = (j + 1) & ~1;
    y = (float)j;
    j -= 2;
    const uint32_t sign = ((uint32_t)(~j & 4)) << 29;
    return sincos_tail(x, y, (j & 2) == 0, sign);
}

/* float64: the reference's *_pd bodies are defective; glibc libm (< 1 ULP)
 * stands in as the correctly-rounded witness. */
int oracle_trig(int op, int t, const void* in, void* out, int64_t n, int64_t s1, int64_t so) {
    if (t == T_F32) {
        switch (op) {
            case G_SIN: UN_LOOP(float, float, oracle_sinf(a))
            case G_COS: UN_LOOP(float, float, oracle_cosf(a))
            case G_LOG: UN_LOOP(float, float, oracle_logf(a))
            case G_EXP: UN_LOOP(float, float, oracle_expf(a))
            /* no reference body (GetTrigOpFast returns NULL, src/atop/ops_trig.cpp:599-777): the reference
             * threads NumPy's loop, i.e. the platform libm; glibc stands in, compared in ULP not bits */
            case G_TAN: UN_LOOP(float, float, tanf(a))
            case G_ASIN: UN_LOOP(float, float, asinf(a))
            case G_ACOS: UN_LOOP(float, float, acosf(a))
            case G_ATAN: UN_LOOP(float, float, atanf(a))
            case G_SINH: UN_LOOP(float, float, sinhf(a))
            case G_COSH: UN_LOOP(float, float, coshf(a))
            case G_TANH: UN_LOOP(float, float, tanhf(a))
            case G_ASINH: UN_LOOP(float, float, asinhf(a))
            case G_ACOSH: UN_LOOP(float, float, acoshf(a))
            case G_ATANH: UN_LOOP(float, float, atanhf(a))
            case G_LOG2: UN_LOOP(float, float, log2f(a))
            case G_LOG10: UN_LOOP(float, float, log10f(a))
            case G_EXP2: UN_LOOP(float, float, exp2f(a))
            case G_EXPM1: UN_LOOP(float, float, expm1f(a))
            case G_LOG1P: UN_LOOP(float, float, log1pf(a))
            case G_CBRT: UN_LOOP(float, float, cbrtf(a))
        }
    } else if (t == T_F64) {
        switch (op) {
            case G_SIN: UN_LOOP(double, double, sin(a))
            case G_COS: UN_LOOP(double, double, cos(a))
            case G_LOG: UN_LOOP(double, double, log(a))
            case G_EXP: UN_LOOP(double, double, exp(a))
            case G_TAN: UN_LOOP(double, double, tan(a))
            case G_ASIN: UN_LOOP(double, double, asin(a))
            case G_ACOS: UN_LOOP(double, double, acos(a))
            case G_ATAN: UN_LOOP(double, double, atan(a))
            case G_SINH: UN_LOOP(double, double, sinh(a))
            case G_COSH: UN_LOOP(double, double, cosh(a))
            case G_TANH: UN_LOOP(double, double, tanh(a))
            case G_ASINH: UN_LOOP(double, double, asinh(a))
            case G_ACOSH: UN_LOOP(double, double, acosh(a))
            case G_ATANH: UN_LOOP(double, double, atanh(a))
            case G_LOG2: UN_LOOP(double, double, log2(a))
            case G_LOG10: UN_LOOP(double, double, log10(a))
            case G_EXP2: UN_LOOP(double, double, exp2(a))
            case G_EXPM1: UN_LOOP(double, double, expm1(a))
            case G_LOG1P: UN_LOOP(double, double, log1p(a))
            case G_CBRT: UN_LOOP(double, double, cbrt(a))
        }
    }
    return -1;
}

/* ------------------------------------------------------------------------------------------
 * reductions: *out = fold(op, in[0..n)) op *start      (start may alias out, or be NULL)
 * follows the reduce branch of AtopBinaryMathFunction with its per-16384-element chunk
 * partials and second pass (src/pnumpy/_pnumpy.cpp:646-731, 359-386) over ReduceMathOpFast /
 * ReduceMathOpFastUpcast (src/atop/ops_binary.cpp:186-344):
 *   - integers wrap in-type, bool ADD = OR, bool MUL = AND  (order independent: exact)
 *   - float32 ADD/MUL: each chunk accumulated in float64, the chunk partial ROUNDED to float32,
 *     partials re-accumulated in float64 and rounded once more -- the reference's intent; its
 *     out-of-bounds read of the horizontal union (ops_binary.cpp:270,309,319)
 *     is not reproduced.
 *   - float64 ADD/MUL: sequential per chunk, then sequential over chunk partials.
 *   - MIN/MAX: NumPy semantics, any NaN -> NaN (the reference runs NumPy's loop per chunk,
 *     ops_binary.cpp:1096-1098).
 * The GPU kernels use a different (tree) association for float sums; tests compare those within
 * the tolerance of SURVEY.md 8c and everything else bit-exactly.
 * ---------------------------------------------------------------------------------------- */
#define CHUNK 16384

#define RED_INT(T, U)                                                          \
    {                                                                          \
        U acc;                                                                 \
        int64_t i = 0;                                                         \
        switch (op) {                                                          \
            case B_ADD: acc = start ? (U) * (const T*)start : 0; for (; i < n; ++i) acc = (U)(acc + (U)AT(T, in, i, s)); break; \
            case B_MUL: acc = start ? (U) * (const T*)start : 1; for (; i < n; ++i) acc = (U)((uint64_t)acc * (uint64_t)(U)AT(T, in, i, s)); break; \
            case B_MIN: { T m = start ? *(const T*)start : AT(T, in, 0, s); for (; i < n; ++i) { T v = AT(T, in, i, s); if (v < m) m = v; } acc = (U)m; } break; \
            case B_MAX: { T m = start ? *(const T*)start : AT(T, in, 0, s); for (; i < n; ++i) { T v = AT(T, in, i, s); if (v > m) m = v; } acc = (U)m; } break; \
            default: return -1;                                                \
        }                                                                      \
        *(T*)out = (T)acc;                                                     \
        return 0;                                                              \
    }

static double chunk_fold_f32(int op, const void* in, int64_t i0, int64_t i1, int64_t s) {
    double acc = (op == B_ADD) ? 0.0 : 1.0;
    for (int64_t i = i0; i < i1; ++i) {
        const double v = (double)AT(float, in, i, s);
        acc = (op == B_ADD) ? acc + v : acc * v;
    }
    return acc;
}

int oracle_reduce(int op, int t, const void* in, void* out, const void* start, int64_t n, int64_t s) {
    if (n <= 0) {
        if (start && start != out) memcpy(out, start, oracle_itemsize(t));
        return 0;
    }
    switch (t) {
        case T_BOOL: {
            unsigned acc;
            if (op == B_ADD || op == B_MAX) { acc = start ? (*(const uint8_t*)start != 0) : 0; for (int64_t i = 0; i < n; ++i) acc |= AT(uint8_t, in, i, s) != 0; }
            else if (op == B_MUL || op == B_MIN) { acc = start ? (*(const uint8_t*)start != 0) : 1; for (int64_t i = 0; i < n; ++i) acc &= AT(uint8_t, in, i, s) != 0; }
            else return -1;
            *(uint8_t*)out = (uint8_t)acc;
            return 0;
        }
        case T_I8: RED_INT(int8_t, uint8_t)
        case T_U8: RED_INT(uint8_t, uint8_t)
        case T_I16: RED_INT(int16_t, uint16_t)
        case T_U16: RED_INT(uint16_t, uint16_t)
        case T_I32: RED_INT(int32_t, uint32_t)
        case T_U32: RED_INT(uint32_t, uint32_t)
        case T_I64: RED_INT(int64_t, uint64_t)
        case T_U64: RED_INT(uint64_t, uint64_t)
        case T_F32: {
            if (op == B_ADD || op == B_MUL) {
                double total = start ? (double)*(const float*)start : (op == B_ADD ? 0.0 : 1.0);
                if (n < 4 * CHUNK) {  /* below the threading threshold: one fold (threads.h:199,1058) */
                    const double p = chunk_fold_f32(op, in, 0, n, s);
                    total = (op == B_ADD) ? total + p : total * p;
                } else {
                    for (int64_t c = 0; c < n; c += CHUNK) {
                        const int64_t e = c + CHUNK < n ? c + CHUNK : n;
                        const float partial = (float)chunk_fold_f32(op, in, c, e, s);
                        total = (op == B_ADD) ? total + (double)partial : total * (double)partial;
                    }
                }
                *(float*)out = (float)total;
                return 0;
            }
            if (op == B_MIN || op == B_MAX) {
                float m = start ? *(const float*)start : AT(float, in, 0, s);
                for (int64_t i = 0; i < n; ++i) {
                    const float v = AT(float, in, i, s);
                    if (op == B_MIN) m = (m < v || m != m) ? m : v; else m = (m > v || m != m) ? m : v;
                }
                *(float*)out = m;
                return 0;
            }
            return -1;
        }
        case T_F64: {
            if (op == B_ADD || op == B_MUL) {
                double total = start ? *(const double*)start : (op == B_ADD ? 0.0 : 1.0);
                for (int64_t c = 0; c < n; c += CHUNK) {
                    const int64_t e = c + CHUNK < n ? c + CHUNK : n;
                    double acc = (op == B_ADD) ? 0.0 : 1.0;
                    for (int64_t i = c; i < e; ++i) acc = (op == B_ADD) ? acc + AT(double, in, i, s) : acc * AT(double, in, i, s);
                    total = (op == B_ADD) ? total + acc : total * acc;
                }
                *(double*)out = total;
                return 0;
            }
            if (op == B_MIN || op == B_MAX) {
                double m = start ? *(const double*)start : AT(double, in, 0, s);
                for (int64_t i = 0; i < n; ++i) {
                    const double v = AT(double, in, i, s);
                    if (op == B_MIN) m = (m < v || m != m) ? m : v; else m = (m > v || m != m) ? m : v;
                }
                *(double*)out = m;
                return 0;
            }
            return -1;
        }
    }
    return -1;
}

/* ------------------------------------------------------------------------------------------
 * argmax (kind 0) / argmin (kind 1): NumPy's PyArray_ArgFunc semantics, which is all the
 * reference provides (AtopArgMinMaxMathFunction forwards, src/pnumpy/_pnumpy.cpp:1105-1118):
 * index of the first extreme element; for floats the first NaN wins.
 * ---------------------------------------------------------------------------------------- */
#define ARG_INT(T)                                                             \
    {                                                                          \
        const T* p = (const T*)in;                                             \
        T m = p[0];                                                            \
        int64_t at = 0;                                                        \
        for (int64_t i = 1; i < n; ++i)                                        \
            if (kind == 0 ? p[i] > m : p[i] < m) { m = p[i]; at = i; }         \
        *idx = at;                                                             \
        return 0;                                                              \
    }
#define ARG_FP(T)                                                              \
    {                                                                          \
        const T* p = (const T*)in;                                             \
        T m = p[0];                                                            \
        int64_t at = 0;                                                        \
        if (m == m)                                                            \
            for (int64_t i = 1; i < n; ++i) {                                  \
                if (p[i] != p[i]) { at = i; break; }                           \
                if (kind == 0 ? p[i] > m : p[i] < m) { m = p[i]; at = i; }     \
            }                                                                  \
        *idx = at;                                                             \
        return 0;                                                              \
    }
int oracle_argminmax(int kind, int t, const void* in, int64_t n, int64_t* idx) {
    if (n <= 0) return -1;
    switch (t) {
        case T_BOOL: case T_U8: ARG_INT(uint8_t)
        case T_I8: ARG_INT(int8_t)
        case T_I16: ARG_INT(int16_t)
        case T_U16: ARG_INT(uint16_t)
        case T_I32: ARG_INT(int32_t)
        case T_U32: ARG_INT(uint32_t)
        case T_I64: ARG_INT(int64_t)
        case T_U64: ARG_INT(uint64_t)
        case T_F32: ARG_FP(float)
        case T_F64: ARG_FP(double)
    }
    return -1;
}

/* ---- dtype conversion (`astype`) ------------------------------------------------------------------
 * Not in the reference yet (first item of its roadmap, doc_src/source/roadmap.rst:11-15); the
 * semantics to match are NumPy's cast loops, which are plain C casts.  This function IS those casts,
 * compiled by the same family of compiler for the same ISA, so values a C cast leaves undefined
 * (float -> int out of range / NaN) come out as x86's cvttss2si / cvttsd2si produce them -- as they do
 * in NumPy.  `volatile` keeps gcc from folding an out-of-range conversion at compile time or
 * vectorising it into an instruction with different saturation.  tests/test_oracle.py pins this
 * against NumPy's astype on random bit patterns for all 121 type pairs. */
/* float -> uint32 as NumPy's (SSE-vectorised) loop does it, observed on NumPy 2.3.5 / x86-64: values below 2^31
 * (and NaN) go through the 32-bit conversion, larger ones through (int32)(x - 2^31) ^ 2^31; what does not fit
 * comes out as 0x80000000 resp. 0.  Integers take the plain cast. */
static inline uint32_t f_to_u32(double x) {
    volatile double v = x;
    if (v >= 2147483648.0) { volatile double w = v - 2147483648.0; return (uint32_t)(int32_t)w ^ 0x80000000u; }
    return (uint32_t)(int32_t)v;
}
#define TO_U32(v) _Generic((v), float: f_to_u32((double)(v)), double: f_to_u32((double)(v)), default: (uint32_t)(v))
#define CONV_ROW(TI, in_is_bool)                                                          \
    switch (to) {                                                                         \
        case T_BOOL: for (i = 0; i < n; ++i) { volatile TI v = ((const TI*)in)[i]; ((uint8_t*)out)[i] = (uint8_t)(v != 0); } return 0; \
        case T_I8:  for (i = 0; i < n; ++i) { volatile TI v = ((const TI*)in)[i]; ((int8_t*)out)[i] = in_is_bool ? (int8_t)(v != 0) : (int8_t)v; } return 0; \
        case T_U8:  for (i = 0; i < n; ++i) { volatile TI v = ((const TI*)in)[i]; ((uint8_t*)out)[i] = in_is_bool ? (uint8_t)(v != 0) : (uint8_t)v; } return 0; \
        case T_I16: for (i = 0; i < n; ++i) { volatile TI v = ((const TI*)in)[i]; ((int16_t*)out)[i] = in_is_bool ? (int16_t)(v != 0) : (int16_t)v; } return 0; \
        case T_U16: for (i = 0; i < n; ++i) { volatile TI v = ((const TI*)in)[i]; ((uint16_t*)out)[i] = in_is_bool ? (uint16_t)(v != 0) : (uint16_t)v; } return 0; \
        case T_I32: for (i = 0; i < n; ++i) { volatile TI v = ((const TI*)in)[i]; ((int32_t*)out)[i] = in_is_bool ? (int32_t)(v != 0) : (int32_t)v; } return 0; \
        case T_U32: for (i = 0; i < n; ++i) { volatile TI v = ((const TI*)in)[i]; ((uint32_t*)out)[i] = in_is_bool ? (uint32_t)(v != 0) : TO_U32(v); } return 0; \
        case T_I64: for (i = 0; i < n; ++i) { volatile TI v = ((const TI*)in)[i]; ((int64_t*)out)[i] = in_is_bool ? (int64_t)(v != 0) : (int64_t)v; } return 0; \
        case T_U64: for (i = 0; i < n; ++i) { volatile TI v = ((const TI*)in)[i]; ((uint64_t*)out)[i] = in_is_bool ? (uint64_t)(v != 0) : (uint64_t)v; } return 0; \
        case T_F32: for (i = 0; i < n; ++i) { volatile TI v = ((const TI*)in)[i]; ((float*)out)[i] = in_is_bool ? (float)(v != 0) : (float)v; } return 0; \
        case T_F64: for (i = 0; i < n; ++i) { volatile TI v = ((const TI*)in)[i]; ((double*)out)[i] = in_is_bool ? (double)(v != 0) : (double)v; } return 0; \
    }                                                                                     \
    return -1;

int oracle_convert(int from, int to, const void* in, void* out, int64_t n) {
    int64_t i;
    switch (from) {
        case T_BOOL: CONV_ROW(uint8_t, 1)
        case T_I8: CONV_ROW(int8_t, 0)
        case T_U8: CONV_ROW(uint8_t, 0)
        case T_I16: CONV_ROW(int16_t, 0)
        case T_U16: CONV_ROW(uint16_t, 0)
        case T_I32: CONV_ROW(int32_t, 0)
        case T_U32: CONV_ROW(uint32_t, 0)
        case T_I64: CONV_ROW(int64_t, 0)
        case T_U64: CONV_ROW(uint64_t, 0)
        case T_F32: CONV_ROW(float, 0)
        case T_F64: CONV_ROW(double, 0)
    }
    return -1;
}
