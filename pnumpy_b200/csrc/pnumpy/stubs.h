// Loop stubs and their lookup tables.
//
// replaces: src/pnumpy/stubs.h (X-macro expansion of one C function per (opcode, atop type)).
// NumPy's ufunc inner-loop signature carries no closure --
//     void loop(char** args, npy_intp const* dimensions, npy_intp const* steps, void* innerloopdata)
// and PyUFunc_ReplaceLoopBySignature only swaps the function pointer -- so, as in the reference,
// every (opcode, type) pair needs its own function that forwards to the category dispatcher with
// (funcop, atype) as constants.  Here the functions are template instantiations and the tables are
// filled by a constexpr index expansion instead of macros; table shapes and index meaning are the
// reference's: g_UFuncGenericLUT[29][22], g_UFuncCompareLUT[6][22], g_UFuncUnaryLUT[35][22],
// g_UFuncTrigLUT[21][22], g_UFuncArgMinMaxLUT[2][22]  (src/pnumpy/stubs.h:79,164,252,368,669).
#pragma once
#include <array>
#include <utility>

static void AtopBinaryMathFunction(char** args, const npy_intp* dimensions, const npy_intp* steps, void* innerloop, int funcop, int atype);
static void AtopCompareMathFunction(char** args, const npy_intp* dimensions, const npy_intp* steps, void* innerloop, int funcop, int atype);
static void AtopUnaryMathFunction(char** args, const npy_intp* dimensions, const npy_intp* steps, void* innerloop, int funcop, int atype);
static void AtopTrigMathFunction(char** args, const npy_intp* dimensions, const npy_intp* steps, void* innerloop, int funcop, int atype);
static int AtopArgMinMaxMathFunction(void* data, npy_intp n, npy_intp* out_index, void* arr, int kind, int atype);

enum StubCategory { STUB_BINARY, STUB_COMPARE, STUB_UNARY, STUB_TRIG };

template <int CAT, int OP, int ATYPE>
static void UFuncStub(char** args, const npy_intp* dimensions, const npy_intp* steps, void* innerloop) {
    if (CAT == STUB_BINARY) AtopBinaryMathFunction(args, dimensions, steps, innerloop, OP, ATYPE);
    else if (CAT == STUB_COMPARE) AtopCompareMathFunction(args, dimensions, steps, innerloop, OP, ATYPE);
    else if (CAT == STUB_UNARY) AtopUnaryMathFunction(args, dimensions, steps, innerloop, OP, ATYPE);
    else AtopTrigMathFunction(args, dimensions, steps, innerloop, OP, ATYPE);
}

template <int KIND, int ATYPE>
static int ArgMinMaxStub(void* data, npy_intp n, npy_intp* out_index, void* arr) {
    return AtopArgMinMaxMathFunction(data, n, out_index, arr, KIND, ATYPE);
}

typedef int (*ArgMinMaxFunc)(void*, npy_intp, npy_intp*, void*);

template <int CAT, int OP, std::size_t... T>
constexpr std::array<PyUFuncGenericFunction, PNCU_ATYPE_LAST> stub_row(std::index_sequence<T...>) {
    return {{&UFuncStub<CAT, OP, (int)T>...}};
}
template <int CAT, std::size_t... OPS>
constexpr std::array<std::array<PyUFuncGenericFunction, PNCU_ATYPE_LAST>, sizeof...(OPS)>
stub_table(std::index_sequence<OPS...>) {
    return {{stub_row<CAT, (int)OPS>(std::make_index_sequence<PNCU_ATYPE_LAST>{})...}};
}
template <int KIND, std::size_t... T>
constexpr std::array<ArgMinMaxFunc, PNCU_ATYPE_LAST> arg_row(std::index_sequence<T...>) {
    return {{&ArgMinMaxStub<KIND, (int)T>...}};
}

static const auto g_UFuncGenericLUT = stub_table<STUB_BINARY>(std::make_index_sequence<PNCU_BINARY_LAST>{});
static const auto g_UFuncCompareLUT = stub_table<STUB_COMPARE>(std::make_index_sequence<PNCU_CMP_LAST>{});
static const auto g_UFuncUnaryLUT = stub_table<STUB_UNARY>(std::make_index_sequence<PNCU_UNARY_LAST>{});
static const auto g_UFuncTrigLUT = stub_table<STUB_TRIG>(std::make_index_sequence<PNCU_TRIG_LAST>{});
static const std::array<std::array<ArgMinMaxFunc, PNCU_ATYPE_LAST>, 2> g_UFuncArgMinMaxLUT = {{
    arg_row<0>(std::make_index_sequence<PNCU_ATYPE_LAST>{}),
    arg_row<1>(std::make_index_sequence<PNCU_ATYPE_LAST>{}),
}};
