// Fused multiply-add chain: out[i] = (a[i] * b[i]) + c[i]   (pncu_GetMulAddFast)
//
// SURVEY.md 8(f) row f-2.  The reference has no body for this: `a*b + c` is two ufunc calls
// (multiply -> a temporary, add), 24 bytes of traffic per float32 element where the fused pass
// moves 16.  The arithmetic is NOT an FMA: the product is rounded to T, then the sum is rounded
// to T -- exactly what np.add(np.multiply(a, b), c) stores -- so the fused result is
// bit-identical to the two-call chain through SimpleMathOpFast (src/atop/ops_binary.cpp:398-571;
// MUL and ADD bodies) and to NumPy.  Integers wrap.
//
// Same shape as ew2_vec_kernel (ew.cuh): full grid, 256-thread CTAs, one 256-bit vector per
// operand per thread, head/tail elements by block 0.  `out` may alias any input exactly.
#include "ew.cuh"
#include "functors.cuh"

using namespace pncu;

namespace {

template <class T> __device__ __forceinline__ T muladd2(T a, T b, T c) { return op_add<T>(op_mul<T>(a, b), c); }

template <class T>
__global__ void __launch_bounds__(EW_BLOCK)
muladd_vec_kernel(const T* a, const T* b, const T* c, T* o, int64_t head, int64_t nvec, int64_t n) {
    constexpr int EPV = EW_VEC_BYTES / (int)sizeof(T);
    const int64_t v = (int64_t)blockIdx.x * EW_BLOCK + threadIdx.x;
    if (v < nvec) {
        const int64_t e0 = head + v * EPV;
        const Pack<T, EPV> ra = ld_pack<T, EPV>(a + e0);
        const Pack<T, EPV> rb = ld_pack<T, EPV>(b + e0);
        const Pack<T, EPV> rc = ld_pack<T, EPV>(c + e0);
        Pack<T, EPV> r;
        r.zero();
#pragma unroll
        for (int e = 0; e < EPV; ++e) r.set(e, muladd2<T>(ra.get(e), rb.get(e), rc.get(e)));
        st_pack<T, EPV>(o + e0, r);
    }
    if (blockIdx.x == 0) {
        const int64_t tail0 = head + nvec * EPV;
        const int64_t nscalar = head + (n - tail0);
        for (int64_t i = threadIdx.x; i < nscalar; i += EW_BLOCK) {
            const int64_t idx = i < head ? i : tail0 + (i - head);
            o[idx] = muladd2<T>(a[idx], b[idx], c[idx]);
        }
    }
}

// operands whose alignments disagree: one element per lane
template <class T>
__global__ void __launch_bounds__(EW_BLOCK)
muladd_elem_kernel(const T* a, const T* b, const T* c, T* o, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * EW_BLOCK + threadIdx.x;
    if (i < n) o[i] = muladd2<T>(a[i], b[i], c[i]);
}

template <class T>
int launch_muladd(const void* in1, const void* in2, const void* in3, void* out, int64_t len, pncu_stream_t stream) {
    if (len <= 0) return 0;
    int rc = ensure_init();
    if (rc) return rc;
    if (!in1 || !in2 || !in3 || !out) { set_error("null operand"); return PNCU_ERR_ARGUMENT; }
    if (((uintptr_t)in1 | (uintptr_t)in2 | (uintptr_t)in3 | (uintptr_t)out) % sizeof(T)) {
        set_error("operand pointer is not a multiple of the item size");
        return PNCU_ERR_UNSUPPORTED;
    }
    if (len > ((int64_t)1 << 38)) { set_error("length %lld exceeds the grid limit", (long long)len); return PNCU_ERR_ARGUMENT; }
    const void* ins[3] = {in1, in2, in3};
    for (int k = 0; k < 3; ++k)
        if ((rc = check_alias(ins[k], sizeof(T), sizeof(T), out, sizeof(T), sizeof(T), len))) return rc;
    cudaStream_t st = as_stream(stream);
    constexpr int EPV = EW_VEC_BYTES / (int)sizeof(T);
    int64_t head = ((EW_VEC_BYTES - (int64_t)((uintptr_t)out % EW_VEC_BYTES)) % EW_VEC_BYTES) / (int64_t)sizeof(T);
    if (head > len) head = len;
    bool ok = true;
    for (int k = 0; k < 3; ++k) ok = ok && ((uintptr_t)ins[k] + head * sizeof(T)) % EW_VEC_BYTES == 0;
    if (ok) {
        const int64_t nvec = (len - head) / EPV;
        muladd_vec_kernel<T><<<full_grid(nvec, EW_BLOCK), EW_BLOCK, 0, st>>>((const T*)in1, (const T*)in2, (const T*)in3,
                                                                           (T*)out, head, nvec, len);
        return after_launch("muladd_vec_kernel");
    }
    muladd_elem_kernel<T><<<full_grid(len, EW_BLOCK), EW_BLOCK, 0, st>>>((const T*)in1, (const T*)in2, (const T*)in3, (T*)out, len);
    return after_launch("muladd_elem_kernel");
}

}  // namespace

extern "C" pncu_any_three_fn pncu_GetMulAddFast(int t) {
    switch (t) {
        case PNCU_INT32: return launch_muladd<int32_t>;
        case PNCU_UINT32: return launch_muladd<uint32_t>;
        case PNCU_INT64: return launch_muladd<int64_t>;
        case PNCU_UINT64: return launch_muladd<uint64_t>;
        case PNCU_FLOAT: return launch_muladd<float>;
        case PNCU_DOUBLE: return launch_muladd<double>;
    }
    return NULL;
}
