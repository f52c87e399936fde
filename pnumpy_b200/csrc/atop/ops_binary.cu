// Binary element-wise getters: pncu_GetSimpleMathOpFast, pncu_GetIntTrueDivide.
// replaces: GetSimpleMathOpFast, src/atop/ops_binary.cpp:871-1058.
//
// Out-type rules are the reference's verbatim (they decide which NumPy loops get hooked,
// src/pnumpy/_pnumpy.cpp:1349); kernel coverage is a superset of the reference's fast
// pointers: every (op, type) the reference hooks has a GPU body.
#include "ew.cuh"
#include "functors.cuh"

using namespace pncu;

namespace {

template <template <class> class F>
pncu_any_two_fn pick_int(int t) {
    switch (t) {
        case PNCU_INT8: return launch_binary<F<int8_t>>;
        case PNCU_UINT8: return launch_binary<F<uint8_t>>;
        case PNCU_INT16: return launch_binary<F<int16_t>>;
        case PNCU_UINT16: return launch_binary<F<uint16_t>>;
        case PNCU_INT32: return launch_binary<F<int32_t>>;
        case PNCU_UINT32: return launch_binary<F<uint32_t>>;
        case PNCU_INT64: return launch_binary<F<int64_t>>;
        case PNCU_UINT64: return launch_binary<F<uint64_t>>;
    }
    return NULL;
}
template <template <class> class F>
pncu_any_two_fn pick_num(int t) {
    switch (t) {
        case PNCU_FLOAT: return launch_binary<F<float>>;
        case PNCU_DOUBLE: return launch_binary<F<double>>;
    }
    return pick_int<F>(t);
}
// bitwise ops depend on the width only
template <template <class> class F>
pncu_any_two_fn pick_bits(int t) {
    switch (t) {
        case PNCU_INT8: case PNCU_UINT8: return launch_binary<F<uint8_t>>;
        case PNCU_INT16: case PNCU_UINT16: return launch_binary<F<uint16_t>>;
        case PNCU_INT32: case PNCU_UINT32: return launch_binary<F<uint32_t>>;
        case PNCU_INT64: case PNCU_UINT64: return launch_binary<F<uint64_t>>;
    }
    return NULL;
}

}  // namespace

extern "C" pncu_any_two_fn pncu_GetSimpleMathOpFast(int func, int t1, int t2, int* wantedOutType) {
    if (t1 != t2 || !wantedOutType) return NULL;  // the hook layer only registers same-type pairs
    const int t = t1;
    switch (func) {
        case PNCU_ADD:  // ops_binary.cpp:876-889
            *wantedOutType = t;
            if (t == PNCU_BOOL) return launch_binary<BoolOrF>;
            return pick_num<AddF>(t);
        case PNCU_MUL:  // ops_binary.cpp:891-908
            *wantedOutType = t;
            if (t == PNCU_BOOL) return launch_binary<BoolAndF>;
            return pick_num<MulF>(t);
        case PNCU_SUB:  // ops_binary.cpp:910-924 (numpy does not subtract bools)
            if (t == PNCU_BOOL) return NULL;
            *wantedOutType = t;
            return pick_num<SubF>(t);
        case PNCU_DIV:  // ops_binary.cpp:926-938 (floats only; int/int is cast by NumPy)
            if (t <= PNCU_UINT64) return NULL;
            *wantedOutType = PNCU_DOUBLE;
            if (t == PNCU_FLOAT) *wantedOutType = PNCU_FLOAT;
            if (t == PNCU_LONGDOUBLE) *wantedOutType = PNCU_LONGDOUBLE;
            if (t == PNCU_FLOAT) return launch_binary<DivF<float>>;
            if (t == PNCU_DOUBLE) return launch_binary<DivF<double>>;
            return NULL;
        case PNCU_FLOORDIV:  // ops_binary.cpp:940-952: floats only
            if (t <= PNCU_UINT64) return NULL;
            *wantedOutType = PNCU_DOUBLE;
            if (t == PNCU_FLOAT) *wantedOutType = PNCU_FLOAT;
            if (t == PNCU_LONGDOUBLE) *wantedOutType = PNCU_LONGDOUBLE;
            if (t == PNCU_FLOAT) return launch_binary<FloorDivF<float>>;
            if (t == PNCU_DOUBLE) return launch_binary<FloorDivF<double>>;
            return NULL;
        case PNCU_MIN:  // ops_binary.cpp:954-963
            *wantedOutType = t;
            if (t == PNCU_BOOL) return launch_binary<BoolAndF>;
            return pick_num<MinF>(t);
        case PNCU_MAX:  // ops_binary.cpp:965-974
            *wantedOutType = t;
            if (t == PNCU_BOOL) return launch_binary<BoolOrF>;
            return pick_num<MaxF>(t);
        case PNCU_LOGICAL_AND:  // ops_binary.cpp:976-982
            if (t != PNCU_BOOL) return NULL;
            *wantedOutType = PNCU_BOOL;
            return launch_binary<BoolAndF>;
        case PNCU_LOGICAL_OR:  // ops_binary.cpp:1000-1006
            if (t != PNCU_BOOL) return NULL;
            *wantedOutType = PNCU_BOOL;
            return launch_binary<BoolOrF>;
        case PNCU_BITWISE_AND:  // ops_binary.cpp:984-998
            if (t > PNCU_UINT64) return NULL;
            *wantedOutType = t;
            if (t == PNCU_BOOL) return launch_binary<BoolAndF>;
            return pick_bits<AndF>(t);
        case PNCU_BITWISE_OR:  // ops_binary.cpp:1008-1022
            if (t > PNCU_UINT64) return NULL;
            *wantedOutType = t;
            if (t == PNCU_BOOL) return launch_binary<BoolOrF>;
            return pick_bits<OrF>(t);
        case PNCU_BITWISE_XOR:  // ops_binary.cpp:1024-1038
            if (t > PNCU_UINT64) return NULL;
            *wantedOutType = t;
            if (t == PNCU_BOOL) return launch_binary<BoolXorF>;
            return pick_bits<XorF>(t);
        case PNCU_BITWISE_LSHIFT:  // ops_binary.cpp:1040-1045 (hooked, NULL body in the reference)
            if (t <= PNCU_BOOL || t > PNCU_UINT64) return NULL;
            *wantedOutType = t;
            return pick_int<LShiftF>(t);
        case PNCU_BITWISE_RSHIFT:  // ops_binary.cpp:1047-1052
            if (t <= PNCU_BOOL || t > PNCU_UINT64) return NULL;
            *wantedOutType = t;
            return pick_int<RShiftF>(t);
    }
    return NULL;
}

extern "C" pncu_any_two_fn pncu_GetIntTrueDivide(int t) {
    switch (t) {
        case PNCU_INT8: return launch_binary<IntTrueDivF<int8_t>>;
        case PNCU_UINT8: return launch_binary<IntTrueDivF<uint8_t>>;
        case PNCU_INT16: return launch_binary<IntTrueDivF<int16_t>>;
        case PNCU_UINT16: return launch_binary<IntTrueDivF<uint16_t>>;
        case PNCU_INT32: return launch_binary<IntTrueDivF<int32_t>>;
        case PNCU_UINT32: return launch_binary<IntTrueDivF<uint32_t>>;
        case PNCU_INT64: return launch_binary<IntTrueDivF<int64_t>>;
        case PNCU_UINT64: return launch_binary<IntTrueDivF<uint64_t>>;
    }
    return NULL;
}
