// Comparison getters: pncu_GetComparisonOpFast.
// replaces: GetComparisonOpFast, src/atop/ops_compare.cpp:1266-1446 and the Compare*
// loop bodies (:400-1255).  Output is always one bool byte (0/1) per element.
#include "ew.cuh"
#include "functors.cuh"

using namespace pncu;

namespace {

template <int OP>
pncu_any_two_fn pick_cmp(int t) {
    switch (t) {
        case PNCU_BOOL: return launch_binary<CmpBoolF<OP>>;
        case PNCU_INT8: return launch_binary<CmpF<int8_t, OP>>;
        case PNCU_UINT8: return launch_binary<CmpF<uint8_t, OP>>;
        case PNCU_INT16: return launch_binary<CmpF<int16_t, OP>>;
        case PNCU_UINT16: return launch_binary<CmpF<uint16_t, OP>>;
        case PNCU_INT32: return launch_binary<CmpF<int32_t, OP>>;
        case PNCU_UINT32: return launch_binary<CmpF<uint32_t, OP>>;
        case PNCU_INT64: return launch_binary<CmpF<int64_t, OP>>;
        case PNCU_UINT64: return launch_binary<CmpF<uint64_t, OP>>;
        case PNCU_FLOAT: return launch_binary<CmpF<float, OP>>;
        case PNCU_DOUBLE: return launch_binary<CmpF<double, OP>>;
    }
    return NULL;
}

}  // namespace

extern "C" pncu_any_two_fn pncu_GetComparisonOpFast(int func, int t1, int t2, int* wantedOutType) {
    // Mixed int64/uint64 pairs have special bodies in the reference (:1343-1390) but its
    // hook layer only ever registers same-type pairs (src/pnumpy/_pnumpy.cpp:1399-1403).
    if (t1 != t2 || !wantedOutType) return NULL;
    *wantedOutType = PNCU_BOOL;  // ops_compare.cpp:1282: always, even when no body exists
    switch (func) {
        case PNCU_CMP_EQ: return pick_cmp<PNCU_CMP_EQ>(t1);
        case PNCU_CMP_NE: return pick_cmp<PNCU_CMP_NE>(t1);
        case PNCU_CMP_LT: return pick_cmp<PNCU_CMP_LT>(t1);
        case PNCU_CMP_GT: return pick_cmp<PNCU_CMP_GT>(t1);
        case PNCU_CMP_LTE: return pick_cmp<PNCU_CMP_LTE>(t1);
        case PNCU_CMP_GTE: return pick_cmp<PNCU_CMP_GTE>(t1);
    }
    return NULL;
}
