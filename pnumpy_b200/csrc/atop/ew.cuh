// Element-wise streaming kernels (the GPU form of the reference's SimpleMathOpFast /
// UnaryOpFast / Compare* loop bodies: src/atop/ops_binary.cpp:398-793,
// src/atop/ops_unary.cpp:327-376, src/atop/ops_compare.cpp:400-1255).
//
// Design (HBM-bound, no reuse, so no shared memory and no tensor cores).  The shape was chosen
// from the sweep in tools/stream_tune.cu on a B200 (profiles/r01_stream_tune.md): for a 2R+1W
// stream, 256-bit accesses, one vector per thread, 256-thread CTAs and a FULL grid (one tile per
// CTA, no grid-stride loop) reached 7.1 TB/s, against 6.2 TB/s for a persistent SMs x k grid.
//   * the unit of work is a 32-byte "wide vector" of the WIDEST operand type; a lane moves one
//     wide vector per operand with one 256-bit ld.global.v8 / st.global.v8 (sm_100 only); the
//     narrower operand (bool result of a compare, int32 input of an int->f64 divide) moves
//     32*narrow/wide bytes per lane with one narrower access.  Consecutive lanes touch
//     consecutive vectors, so every warp-level access is fully coalesced.
//   * one CTA = one tile of EW_BLOCK vectors; the hardware CTA scheduler balances the tail.
//   * head/tail elements that do not fill an aligned wide vector are done by block 0 in the same
//     launch (no second kernel).
//   * `out` may alias an input exactly (in-place): every element is read and written by the same
//     thread, loads precede stores (volatile asm keeps program order), nothing is __restrict__.
#pragma once
#include "common.cuh"
#include "functors.cuh"

namespace pncu {

template <int SIZE> struct Bits;
template <> struct Bits<1> { typedef uint8_t type; };
template <> struct Bits<2> { typedef uint16_t type; };
template <> struct Bits<4> { typedef uint32_t type; };
template <> struct Bits<8> { typedef uint64_t type; };

template <class T> __device__ __forceinline__ typename Bits<sizeof(T)>::type to_bits(T v) {
    union { T t; typename Bits<sizeof(T)>::type b; } c;
    c.t = v;
    return c.b;
}
template <class T> __device__ __forceinline__ T from_bits(typename Bits<sizeof(T)>::type b) {
    union { T t; typename Bits<sizeof(T)>::type b; } c;
    c.b = b;
    return c.t;
}

// ---- register image of one lane's share of a vector ---------------------------------------
// N elements of T held in 32-bit words.  Elements are extracted / inserted with shifts on
// compile-time indices so everything stays in registers (a union of byte arrays would be
// demoted to local memory by the compiler for 1- and 2-byte types).
template <class T, int N>
struct Pack {
    static constexpr int BYTES = (int)sizeof(T) * N;
    static constexpr int WORDS = (BYTES + 3) / 4;
    uint32_t w[WORDS];

    __device__ __forceinline__ T get(int e) const {
        if constexpr (sizeof(T) == 1) return from_bits<T>((uint8_t)(w[e >> 2] >> (8 * (e & 3))));
        else if constexpr (sizeof(T) == 2) return from_bits<T>((uint16_t)(w[e >> 1] >> (16 * (e & 1))));
        else if constexpr (sizeof(T) == 4) return from_bits<T>(w[e]);
        else return from_bits<T>(((uint64_t)w[2 * e + 1] << 32) | (uint64_t)w[2 * e]);
    }
    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int i = 0; i < WORDS; ++i) w[i] = 0u;
    }
    // insert element e; for sub-word T the pack must have been zero()ed first
    __device__ __forceinline__ void set(int e, T v) {
        if constexpr (sizeof(T) == 1) w[e >> 2] |= (uint32_t)to_bits<T>(v) << (8 * (e & 3));
        else if constexpr (sizeof(T) == 2) w[e >> 1] |= (uint32_t)to_bits<T>(v) << (16 * (e & 1));
        else if constexpr (sizeof(T) == 4) w[e] = to_bits<T>(v);
        else {
            const uint64_t b = to_bits<T>(v);
            w[2 * e] = (uint32_t)b;
            w[2 * e + 1] = (uint32_t)(b >> 32);
        }
    }
};

// one global access of BYTES in {1,2,4,8,16,32}; volatile asm keeps loads before stores
template <int BYTES>
__device__ __forceinline__ void ld_words(const void* p, uint32_t* w) {
    if constexpr (BYTES == 32)
        asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                     : "l"(p) : "memory");
    else if constexpr (BYTES == 16)
        asm volatile("ld.global.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "l"(p) : "memory");
    else if constexpr (BYTES == 8)
        asm volatile("ld.global.v2.b32 {%0,%1}, [%2];" : "=r"(w[0]), "=r"(w[1]) : "l"(p) : "memory");
    else if constexpr (BYTES == 4)
        asm volatile("ld.global.b32 %0, [%1];" : "=r"(w[0]) : "l"(p) : "memory");
    else if constexpr (BYTES == 2)
        asm volatile("{ .reg .b16 t; ld.global.b16 t, [%1]; cvt.u32.u16 %0, t; }" : "=r"(w[0]) : "l"(p) : "memory");
    else
        asm volatile("ld.global.u8 %0, [%1];" : "=r"(w[0]) : "l"(p) : "memory");
}
template <int BYTES>
__device__ __forceinline__ void st_words(void* p, const uint32_t* w) {
    if constexpr (BYTES == 32)
        asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                     :: "l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
    else if constexpr (BYTES == 16)
        asm volatile("st.global.v4.b32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
    else if constexpr (BYTES == 8)
        asm volatile("st.global.v2.b32 [%0], {%1,%2};" :: "l"(p), "r"(w[0]), "r"(w[1]) : "memory");
    else if constexpr (BYTES == 4)
        asm volatile("st.global.b32 [%0], %1;" :: "l"(p), "r"(w[0]) : "memory");
    else if constexpr (BYTES == 2)
        asm volatile("{ .reg .b16 t; cvt.u16.u32 t, %1; st.global.b16 [%0], t; }" :: "l"(p), "r"(w[0]) : "memory");
    else
        asm volatile("st.global.u8 [%0], %1;" :: "l"(p), "r"(w[0]) : "memory");
}

template <class T, int N>
__device__ __forceinline__ Pack<T, N> ld_pack(const T* p) {
    Pack<T, N> r;
    ld_words<Pack<T, N>::BYTES>(p, r.w);
    return r;
}
template <class T, int N>
__device__ __forceinline__ void st_pack(T* p, const Pack<T, N>& v) {
    st_words<Pack<T, N>::BYTES>(p, v.w);
}

// ---- operands whose alignment disagrees with the output's ----------------------------------------------
// The vector kernels store aligned wide vectors of `out`.  A contiguous input that sits `m` bytes past its own
// vector alignment at the same element (a[1:] next to a[:-1], any view cut at an odd element) would need a
// misaligned 256-bit load.  Instead every lane loads the ALIGNED vector that contains its first byte, takes the
// next aligned vector from the NEXT LANE's registers (one shuffle per word; lane 31 and the last active lane
// load theirs directly) and extracts bytes [m, m + BYTES) of the pair with funnel shifts.  Same bytes from HBM,
// still one fully coalesced 256-bit load per lane, and the arithmetic sees exactly the elements it would have
// seen -- results cannot differ.  (The reference's AVX2 loops simply use unaligned loads, loadu,
// src/atop/ops_binary.cpp:600-616.)  The launcher guarantees that both aligned vectors of every active lane lie
// inside the operand (it widens the scalar head / tail by one vector), so nothing outside the array is read.
template <int W, int WS>
struct FunnelExtract {
    static __device__ __forceinline__ void run(const uint32_t (&A)[W], const uint32_t (&N)[W], unsigned ws, unsigned bs, uint32_t* out) {
        if (ws == WS) {
#pragma unroll
            for (int i = 0; i < W; ++i) {
                const uint32_t lo = (i + WS < W) ? A[(i + WS) % W] : N[(i + WS) % W];
                const uint32_t hi = (i + WS + 1 < W) ? A[(i + WS + 1) % W] : N[(i + WS + 1) % W];
                out[i] = __funnelshift_r(lo, hi, bs);
            }
        } else {
            FunnelExtract<W, WS + 1>::run(A, N, ws, bs, out);
        }
    }
};
template <int W>
struct FunnelExtract<W, W> {
    static __device__ __forceinline__ void run(const uint32_t (&)[W], const uint32_t (&)[W], unsigned, unsigned, uint32_t*) {}
};

// `p`: this lane's first element (element-aligned); `m` = (address of p) mod BYTES, warp-uniform; `active`: this
// lane has a vector; `next_active`: so has the next lane.  Must be called by all 32 lanes of the warp.
template <class T, int N>
__device__ __forceinline__ Pack<T, N> ld_pack_realigned(const T* p, unsigned m, bool active, bool next_active) {
    constexpr int BYTES = Pack<T, N>::BYTES;
    static_assert(BYTES % 4 == 0, "vectors are whole words");
    constexpr int W = BYTES / 4;
    Pack<T, N> r;
    if (m == 0) {   // this operand agrees with the output after all (uniform branch)
        if (active) r = ld_pack<T, N>(p); else r.zero();
        return r;
    }
    const char* base = reinterpret_cast<const char*>(p) - m;
    uint32_t A[W], Nx[W];
#pragma unroll
    for (int w = 0; w < W; ++w) A[w] = 0u;
    if (active) ld_words<BYTES>(base, A);
#pragma unroll
    for (int w = 0; w < W; ++w) Nx[w] = __shfl_down_sync(0xffffffffu, A[w], 1);
    if (active && ((threadIdx.x & 31) == 31 || !next_active)) ld_words<BYTES>(base + BYTES, Nx);
    FunnelExtract<W, 0>::run(A, Nx, m >> 2, (m & 3u) * 8u, r.w);
    return r;
}

__host__ __device__ constexpr int cmax(int a, int b) { return a > b ? a : b; }

constexpr int EW_BLOCK = 256;   // threads per CTA == vectors per tile

enum Layout { LAY_VV = 0, LAY_SV = 1, LAY_VS = 2 };


constexpr int EW_VEC_BYTES = 32;

// elements per wide vector for an (In, Out) pair
template <class In, class Out>
struct Epv { static constexpr int value = EW_VEC_BYTES / cmax((int)sizeof(In), (int)sizeof(Out)); };

// ---- binary ----------------------------------------------------------------------
// F: struct { typedef In; typedef Out; static __device__ Out apply(In, In); }
template <class F, int LAYOUT, bool REALIGN>
__global__ void __launch_bounds__(EW_BLOCK)
ew2_vec_kernel(const typename F::In* a, const typename F::In* b, typename F::Out* o,
               int64_t head, int64_t nvec, int64_t n, unsigned ma, unsigned mb) {
    typedef typename F::In In;
    typedef typename F::Out Out;
    constexpr int EPV = Epv<In, Out>::value;

    In sa = In(), sb = In();
    if (LAYOUT == LAY_SV) sa = a[0];
    if (LAYOUT == LAY_VS) sb = b[0];

    const int64_t v = (int64_t)blockIdx.x * EW_BLOCK + threadIdx.x;
    const bool active = v < nvec;
    if (REALIGN || active) {
        const int64_t e0 = head + v * EPV;
        Pack<In, EPV> ra, rb;
        if constexpr (REALIGN) {   // every lane takes part in the shuffles, active or not
            if (LAYOUT != LAY_SV) ra = ld_pack_realigned<In, EPV>(a + e0, ma, active, v + 1 < nvec);
            if (LAYOUT != LAY_VS) rb = ld_pack_realigned<In, EPV>(b + e0, mb, active, v + 1 < nvec);
        } else {
            if (LAYOUT != LAY_SV) ra = ld_pack<In, EPV>(a + e0);
            if (LAYOUT != LAY_VS) rb = ld_pack<In, EPV>(b + e0);
        }
        if (active) {
            Pack<Out, EPV> r;
            if constexpr (WordOp<F>::enabled && sizeof(In) == sizeof(Out)) {
                // 1- and 2-byte types: whole words through the SIMD-in-a-word form (functors.cuh)
                const unsigned wa = splat_word<In>(sa), wb = splat_word<In>(sb);
#pragma unroll
                for (int w = 0; w < Pack<Out, EPV>::WORDS; ++w)
                    r.w[w] = WordOp<F>::word(LAYOUT == LAY_SV ? wa : ra.w[w], LAYOUT == LAY_VS ? wb : rb.w[w]);
            } else {
                r.zero();
#pragma unroll
                for (int e = 0; e < EPV; ++e)
                    r.set(e, F::apply(LAYOUT == LAY_SV ? sa : ra.get(e), LAYOUT == LAY_VS ? sb : rb.get(e)));
            }
            st_pack<Out, EPV>(o + e0, r);
        }
    }
    // head + tail scalars: elements before the first / after the last aligned wide vector
    if (blockIdx.x == 0) {
        const int64_t tail0 = head + nvec * EPV;
        const int64_t nscalar = head + (n - tail0);
        for (int64_t i = threadIdx.x; i < nscalar; i += EW_BLOCK) {
            const int64_t idx = i < head ? i : tail0 + (i - head);
            o[idx] = F::apply(LAYOUT == LAY_SV ? sa : a[idx], LAYOUT == LAY_VS ? sb : b[idx]);
        }
    }
}

// generic byte-strided fallback (any stride incl. 0 and negative); one element per lane
template <class F>
__global__ void __launch_bounds__(EW_BLOCK)
ew2_strided_kernel(const char* a, const char* b, char* o, int64_t n,
                   int64_t s1, int64_t s2, int64_t so) {
    typedef typename F::In In;
    typedef typename F::Out Out;
    const int64_t i = (int64_t)blockIdx.x * EW_BLOCK + threadIdx.x;
    if (i < n) {
        const In x = *reinterpret_cast<const In*>(a + i * s1);
        const In y = *reinterpret_cast<const In*>(b + i * s2);
        *reinterpret_cast<Out*>(o + i * so) = F::apply(x, y);
    }
}

// ---- unary -------------------------------------------------------------------------
// F: struct { typedef In; typedef Out; static __device__ Out apply(In); }
// Compute-heavy functors (sin/log/exp: ~25-40 instructions per element) set kUnroll > 1: a thread
// then issues kUnroll 256-bit loads before its first use.  With one load per thread the time a warp
// spends computing is time it has nothing in flight, and 64 warps/SM x 32 B no longer cover the
// HBM latency-bandwidth product (measured: sin at 5.8 TB/s with 1 vector per thread).
#ifndef PNCU_EW1_UNROLL
#define PNCU_EW1_UNROLL 1
#endif
template <class F, class = void> struct unroll_of { static constexpr int value = PNCU_EW1_UNROLL; };
template <class F> struct unroll_of<F, decltype((void)F::kUnroll)> { static constexpr int value = F::kUnroll; };

// Functors with a branch-free FAST body valid on part of the domain declare
//     static constexpr bool kHasFast = true;  static bool ok(In);  static Out fast(In [, const void* table]);
// and make apply() the (out-of-line) general function.  The kernel then tests a thread's whole set of
// inputs once: if every element is in the fast domain (the overwhelmingly common case) the packs run
// the fast body with no per-element branch, so the independent chains interleave; otherwise the
// thread falls back to apply() element by element.
template <class F, class = void> struct has_fast { static constexpr bool value = false; };
template <class F> struct has_fast<F, decltype((void)F::kHasFast)> { static constexpr bool value = F::kHasFast; };
// Functors whose fast body reads a lookup table from shared memory declare
//     static constexpr int kSmemBytes;  static void smem_init(void* smem) (all threads of the CTA);
//     static const void* smem_lane(const void* smem) (this lane's view);  static Out fast(In, const void*)
template <class F, class = void> struct smem_of { static constexpr int value = 0; };
template <class F> struct smem_of<F, decltype((void)F::kSmemBytes)> { static constexpr int value = F::kSmemBytes; };

template <class F, int EPV>
__device__ __forceinline__ bool pack_ok(const Pack<typename F::In, EPV>& ra) {
    bool ok = true;
#pragma unroll
    for (int e = 0; e < EPV; ++e) ok = ok && F::ok(ra.get(e));
    return ok;
}
template <class F, int EPV>
__device__ __forceinline__ Pack<typename F::Out, EPV> fast_pack(const Pack<typename F::In, EPV>& ra, const void* table) {
    typedef typename F::Out Out;
    Pack<Out, EPV> r;
    r.zero();
#pragma unroll
    for (int e = 0; e < EPV; ++e) {
        if constexpr (smem_of<F>::value != 0) r.set(e, F::fast(ra.get(e), table));
        else r.set(e, F::fast(ra.get(e)));
    }
    return r;
}

template <class F, int EPV>
__device__ __forceinline__ Pack<typename F::Out, EPV> apply_pack(const Pack<typename F::In, EPV>& ra) {
    typedef typename F::Out Out;
    Pack<Out, EPV> r;
    if constexpr (WordOp1<F>::enabled && sizeof(typename F::In) == sizeof(Out)) {
#pragma unroll
        for (int w = 0; w < Pack<Out, EPV>::WORDS; ++w) r.w[w] = WordOp1<F>::word(ra.w[w]);
    } else {
        r.zero();
#pragma unroll
        for (int e = 0; e < EPV; ++e) r.set(e, F::apply(ra.get(e)));
    }
    return r;
}

template <class F>
__global__ void __launch_bounds__(EW_BLOCK)
ew1_vec_kernel(const typename F::In* a, typename F::Out* o, int64_t head, int64_t nvec, int64_t n) {
    typedef typename F::In In;
    typedef typename F::Out Out;
    constexpr int EPV = Epv<In, Out>::value;
    constexpr int U = unroll_of<F>::value;
    // the branch-free fast body (table functors keep theirs for ew1_table_kernel: no table here)
    constexpr bool FAST = has_fast<F>::value && smem_of<F>::value == 0;
    const int64_t v0 = (int64_t)blockIdx.x * (EW_BLOCK * U) + threadIdx.x;
    const bool full = v0 + (int64_t)(U - 1) * EW_BLOCK < nvec;
    if (full) {
        Pack<In, EPV> ra[U];
#pragma unroll
        for (int u = 0; u < U; ++u) ra[u] = ld_pack<In, EPV>(a + head + (v0 + (int64_t)u * EW_BLOCK) * EPV);
        bool all_fast = FAST;
        if constexpr (FAST) {
#pragma unroll
            for (int u = 0; u < U; ++u) all_fast = all_fast && pack_ok<F, EPV>(ra[u]);
        }
        if (all_fast) {
            if constexpr (FAST) {
#pragma unroll
                for (int u = 0; u < U; ++u)
                    st_pack<Out, EPV>(o + head + (v0 + (int64_t)u * EW_BLOCK) * EPV, fast_pack<F, EPV>(ra[u], nullptr));
            }
        } else {
#pragma unroll
            for (int u = 0; u < U; ++u)
                st_pack<Out, EPV>(o + head + (v0 + (int64_t)u * EW_BLOCK) * EPV, apply_pack<F, EPV>(ra[u]));
        }
    } else {
#pragma unroll 1
        for (int u = 0; u < U; ++u) {
            const int64_t v = v0 + (int64_t)u * EW_BLOCK;
            if (v < nvec) st_pack<Out, EPV>(o + head + v * EPV, apply_pack<F, EPV>(ld_pack<In, EPV>(a + head + v * EPV)));
        }
    }
    if (blockIdx.x == 0) {
        const int64_t tail0 = head + nvec * EPV;
        const int64_t nscalar = head + (n - tail0);
        for (int64_t i = threadIdx.x; i < nscalar; i += EW_BLOCK) {
            const int64_t idx = i < head ? i : tail0 + (i - head);
            o[idx] = F::apply(a[idx]);
        }
    }
}

// The same for an input whose 32-byte alignment disagrees with the output's (np.sin(a[1:])).  Every lane loads the
// ALIGNED input vector that holds the first byte of its output vector and takes the following one from the NEXT LANE
// (one shuffle per word) -- and so that no lane ever needs another warp's registers, a warp produces only 31 output
// vectors: lane 31 loads and hands over, nothing else.  Costs 3 % of the lanes and nothing else (the overlapping
// vector is an L1/L2 hit); a first version that gave lane 31 the next warp's vector through an extra divergent load,
// and a second one through shared memory, added ~10 instructions per element to bodies that are issue-bound
// (sin(a[1:]): 59-62 % of 8 TB/s; ncu: +29 % instructions, 3 CTAs per SM).  All U vectors of a thread are loaded
// before the first shuffle.  The launcher guarantees that aligned vectors 0 .. nvec lie inside the operand.
constexpr int EW_RL_VPW = 31;                              // output vectors per warp
constexpr int EW_RL_VPB = (EW_BLOCK / 32) * EW_RL_VPW;     // per CTA and unroll step

template <class F>
__global__ void __launch_bounds__(EW_BLOCK)
ew1_realign_kernel(const typename F::In* a, typename F::Out* o, int64_t head, int64_t nvec, int64_t n, unsigned ma) {
    typedef typename F::In In;
    typedef typename F::Out Out;
    constexpr int EPV = Epv<In, Out>::value;
    constexpr int U = unroll_of<F>::value;
    constexpr bool FAST = has_fast<F>::value && smem_of<F>::value == 0;
    constexpr int BYTES = Pack<In, EPV>::BYTES;
    static_assert(BYTES % 4 == 0, "vectors are whole words");
    constexpr int W = BYTES / 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t v0 = (int64_t)blockIdx.x * (EW_RL_VPB * U) + warp * EW_RL_VPW + lane;   // lane 31: the next warp's first
    const bool out_lane = lane < EW_RL_VPW;
    const char* base = reinterpret_cast<const char*>(a + head) - ma;   // aligned input vector v starts at base + v BYTES
    uint32_t A[U][W];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int64_t v = v0 + (int64_t)u * EW_RL_VPB;
#pragma unroll
        for (int w = 0; w < W; ++w) A[u][w] = 0u;
        if (v <= nvec && nvec > 0) ld_words<BYTES>(base + v * BYTES, A[u]);   // (nvec == 0: all scalars, `base` means nothing)
    }
    Pack<In, EPV> ra[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        uint32_t Nx[W];
#pragma unroll
        for (int w = 0; w < W; ++w) Nx[w] = __shfl_down_sync(0xffffffffu, A[u][w], 1);
        FunnelExtract<W, 0>::run(A[u], Nx, ma >> 2, (ma & 3u) * 8u, ra[u].w);
    }
    const bool full = out_lane && v0 + (int64_t)(U - 1) * EW_RL_VPB < nvec;
    if (full) {
        bool all_fast = FAST;
        if constexpr (FAST) {
#pragma unroll
            for (int u = 0; u < U; ++u) all_fast = all_fast && pack_ok<F, EPV>(ra[u]);
        }
        if (all_fast) {
            if constexpr (FAST) {
#pragma unroll
                for (int u = 0; u < U; ++u)
                    st_pack<Out, EPV>(o + head + (v0 + (int64_t)u * EW_RL_VPB) * EPV, fast_pack<F, EPV>(ra[u], nullptr));
            }
        } else {
#pragma unroll
            for (int u = 0; u < U; ++u)
                st_pack<Out, EPV>(o + head + (v0 + (int64_t)u * EW_RL_VPB) * EPV, apply_pack<F, EPV>(ra[u]));
        }
    } else if (out_lane) {
        // (unrolled: a dynamic index would put ra[] in local memory for the whole kernel)
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t v = v0 + (int64_t)u * EW_RL_VPB;
            if (v < nvec) st_pack<Out, EPV>(o + head + v * EPV, apply_pack<F, EPV>(ra[u]));
        }
    }
    if (blockIdx.x == 0) {
        const int64_t tail0 = head + nvec * EPV;
        const int64_t nscalar = head + (n - tail0);
        for (int64_t i = threadIdx.x; i < nscalar; i += EW_BLOCK) {
            const int64_t idx = i < head ? i : tail0 + (i - head);
            o[idx] = F::apply(a[idx]);
        }
    }
}

// Variant for functors with a shared-memory lookup table (float64 log): a CTA builds the table once
// and then streams EW_TABLE_TILES tiles, so the set-up (16 KiB of shared-memory stores) is paid once
// per 256 KiB of traffic.  The first tile's loads are issued BEFORE the table is built and every
// later tile's loads before the previous tile is computed, so neither the set-up nor the arithmetic
// leaves the thread with nothing in flight.  The grid is still thousands of CTAs: the hardware
// scheduler balances it.
constexpr int EW_TABLE_TILES = 8;

template <class F>
__global__ void __launch_bounds__(EW_BLOCK)
ew1_table_kernel(const typename F::In* a, typename F::Out* o, int64_t head, int64_t nvec, int64_t n) {
    typedef typename F::In In;
    typedef typename F::Out Out;
    constexpr int EPV = Epv<In, Out>::value;
    constexpr int U = unroll_of<F>::value;
    static_assert(has_fast<F>::value, "a table functor needs ok()/fast()");
    __shared__ __align__(16) unsigned char smem[smem_of<F>::value];
    // a tile is "full" when all U vectors of this thread exist; ragged tiles take the slow loop below
    const int64_t vbase = (int64_t)blockIdx.x * EW_TABLE_TILES * (EW_BLOCK * U) + threadIdx.x;
    Pack<In, EPV> cur[U];
    bool cur_full = vbase + (int64_t)(U - 1) * EW_BLOCK < nvec;
    if (cur_full) {
#pragma unroll
        for (int u = 0; u < U; ++u) cur[u] = ld_pack<In, EPV>(a + head + (vbase + (int64_t)u * EW_BLOCK) * EPV);
    }
    F::smem_init(smem);
    __syncthreads();
    const void* table = F::smem_lane(smem);
#pragma unroll 1
    for (int t = 0; t < EW_TABLE_TILES; ++t) {
        const int64_t v0 = vbase + (int64_t)t * (EW_BLOCK * U);
        if (v0 >= nvec) break;
        const int64_t v1 = v0 + (EW_BLOCK * U);
        Pack<In, EPV> nxt[U];
        const bool nxt_full = t + 1 < EW_TABLE_TILES && v1 + (int64_t)(U - 1) * EW_BLOCK < nvec;
        if (nxt_full) {
#pragma unroll
            for (int u = 0; u < U; ++u) nxt[u] = ld_pack<In, EPV>(a + head + (v1 + (int64_t)u * EW_BLOCK) * EPV);
        }
        if (cur_full) {
            bool all_fast = true;
#pragma unroll
            for (int u = 0; u < U; ++u) all_fast = all_fast && pack_ok<F, EPV>(cur[u]);
            if (all_fast) {
#pragma unroll
                for (int u = 0; u < U; ++u)
                    st_pack<Out, EPV>(o + head + (v0 + (int64_t)u * EW_BLOCK) * EPV, fast_pack<F, EPV>(cur[u], table));
            } else {
#pragma unroll
                for (int u = 0; u < U; ++u)
                    st_pack<Out, EPV>(o + head + (v0 + (int64_t)u * EW_BLOCK) * EPV, apply_pack<F, EPV>(cur[u]));
            }
        } else {
#pragma unroll 1
            for (int u = 0; u < U; ++u) {
                const int64_t v = v0 + (int64_t)u * EW_BLOCK;
                if (v < nvec) {
                    const Pack<In, EPV> ra = ld_pack<In, EPV>(a + head + v * EPV);
                    st_pack<Out, EPV>(o + head + v * EPV, apply_pack<F, EPV>(ra));
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) cur[u] = nxt[u];
        cur_full = nxt_full;
    }
    if (blockIdx.x == 0) {
        const int64_t tail0 = head + nvec * EPV;
        const int64_t nscalar = head + (n - tail0);
        for (int64_t i = threadIdx.x; i < nscalar; i += EW_BLOCK) {
            const int64_t idx = i < head ? i : tail0 + (i - head);
            o[idx] = F::apply(a[idx]);
        }
    }
}

template <class F>
__global__ void __launch_bounds__(EW_BLOCK)
ew1_strided_kernel(const char* a, char* o, int64_t n, int64_t s1, int64_t so) {
    typedef typename F::In In;
    typedef typename F::Out Out;
    const int64_t i = (int64_t)blockIdx.x * EW_BLOCK + threadIdx.x;
    if (i < n) {
        const In x = *reinterpret_cast<const In*>(a + i * s1);
        *reinterpret_cast<Out*>(o + i * so) = F::apply(x);
    }
}

// ---- host side: layout classification + launch -------------------------------------------

inline bool ranges_overlap(const char* p, int64_t pbytes, const char* q, int64_t qbytes) {
    return p < q + qbytes && q < p + pbytes;
}
// byte extent [lo, lo+bytes) of `len` elements of `size` bytes at byte stride `s`
inline void extent(const void* p, int64_t len, int64_t s, int64_t size, const char** lo, int64_t* bytes) {
    const char* c = (const char*)p;
    int64_t span = (len - 1) * s;
    if (span >= 0) { *lo = c; *bytes = span + size; }
    else { *lo = c + span; *bytes = -span + size; }
}

// 0 ok, PNCU_ERR_OVERLAP if `out` overlaps `in` other than exact element-wise aliasing
// (the reference drops to its scalar loop for near overlaps, src/pnumpy/_pnumpy.cpp:743-745;
// a parallel kernel cannot honour sequential semantics, so the hook layer is told to decline)
inline int check_alias(const void* in, int64_t sin, int64_t insize, const void* out, int64_t sout,
                       int64_t outsize, int64_t len) {
    const char *ilo, *olo;
    int64_t ib, ob;
    extent(in, len, sin, insize, &ilo, &ib);
    extent(out, len, sout, outsize, &olo, &ob);
    if (!ranges_overlap(ilo, ib, olo, ob)) return 0;
    if (in == out && sin == sout && insize == outsize) return 0;
    set_error("output overlaps an input without aliasing it exactly (in=%p stride %lld, out=%p stride %lld, len %lld)",
              in, (long long)sin, out, (long long)sout, (long long)len);
    return PNCU_ERR_OVERLAP;
}

inline unsigned full_grid(int64_t items, int per_cta) {
    int64_t g = (items + per_cta - 1) / per_cta;
    return (unsigned)(g < 1 ? 1 : g);
}

template <class F>
int launch_binary(const void* in1, const void* in2, void* out, int64_t len, int64_t s1, int64_t s2,
                  int64_t so, pncu_stream_t stream) {
    typedef typename F::In In;
    typedef typename F::Out Out;
    if (len <= 0) return 0;
    int rc = ensure_init();
    if (rc) return rc;
    if (!in1 || !in2 || !out) { set_error("null operand"); return PNCU_ERR_ARGUMENT; }
    if (((uintptr_t)in1 | (uintptr_t)in2 | (uint64_t)s1 | (uint64_t)s2) % sizeof(In) ||
        ((uintptr_t)out | (uint64_t)so) % sizeof(Out)) {
        set_error("operand pointer or stride is not a multiple of the item size");
        return PNCU_ERR_UNSUPPORTED;
    }
    if (len > ((int64_t)1 << 38)) { set_error("length %lld exceeds the grid limit", (long long)len); return PNCU_ERR_ARGUMENT; }
    if ((rc = check_alias(in1, s1, sizeof(In), out, so, sizeof(Out), len))) return rc;
    if ((rc = check_alias(in2, s2, sizeof(In), out, so, sizeof(Out), len))) return rc;
    cudaStream_t st = as_stream(stream);

    constexpr int EPV = Epv<In, Out>::value;
    constexpr int IN_VB = EPV * (int)sizeof(In);    // bytes one lane moves per input vector
    constexpr int OUT_VB = EPV * (int)sizeof(Out);
    const bool c1 = s1 == (int64_t)sizeof(In), c2 = s2 == (int64_t)sizeof(In);
    const bool z1 = s1 == 0, z2 = s2 == 0;
    if (so == (int64_t)sizeof(Out) && ((c1 && c2) || (z1 && c2) || (c1 && z2))) {
        // elements to peel so that `out` reaches its vector alignment
        int64_t head = ((OUT_VB - (int64_t)((uintptr_t)out % OUT_VB)) % OUT_VB) / (int64_t)sizeof(Out);
        if (head > len) head = len;
        // where does each input stand at that element?  0 = agrees with `out` (plain aligned vector loads)
        unsigned ma = z1 ? 0u : (unsigned)(((uintptr_t)in1 + head * sizeof(In)) % IN_VB);
        unsigned mb = z2 ? 0u : (unsigned)(((uintptr_t)in2 + head * sizeof(In)) % IN_VB);
        int64_t nvec = (len - head) / EPV;
        const bool realign = (ma | mb) != 0 && IN_VB % 4 == 0;
        if (realign) {
            // a realigning lane reads the aligned vector that contains its first byte and the one after it: keep
            // both inside the operand by leaving one more vector to the scalar head (if the first aligned vector
            // would start before the array) and one to the scalar tail
            if ((int64_t)(head * sizeof(In)) < (int64_t)(ma > mb ? ma : mb)) head += EPV;
            if (head > len) head = len;
            nvec = (len - head) / EPV - 1;
            if (nvec < 0) nvec = 0;
        }
        if (realign || (ma | mb) == 0) {
            const unsigned grid = full_grid(nvec, EW_BLOCK);
            const In* a = (const In*)in1;
            const In* b = (const In*)in2;
            Out* o = (Out*)out;
            if (realign) {
                if constexpr (IN_VB % 4 == 0) {
                    if (c1 && c2) ew2_vec_kernel<F, LAY_VV, true><<<grid, EW_BLOCK, 0, st>>>(a, b, o, head, nvec, len, ma, mb);
                    else if (z1) ew2_vec_kernel<F, LAY_SV, true><<<grid, EW_BLOCK, 0, st>>>(a, b, o, head, nvec, len, ma, mb);
                    else ew2_vec_kernel<F, LAY_VS, true><<<grid, EW_BLOCK, 0, st>>>(a, b, o, head, nvec, len, ma, mb);
                }
            } else if (c1 && c2)
                ew2_vec_kernel<F, LAY_VV, false><<<grid, EW_BLOCK, 0, st>>>(a, b, o, head, nvec, len, 0u, 0u);
            else if (z1)
                ew2_vec_kernel<F, LAY_SV, false><<<grid, EW_BLOCK, 0, st>>>(a, b, o, head, nvec, len, 0u, 0u);
            else
                ew2_vec_kernel<F, LAY_VS, false><<<grid, EW_BLOCK, 0, st>>>(a, b, o, head, nvec, len, 0u, 0u);
            return after_launch("ew2_vec_kernel");
        }
    }
    ew2_strided_kernel<F><<<full_grid(len, EW_BLOCK), EW_BLOCK, 0, st>>>((const char*)in1, (const char*)in2,
                                                                         (char*)out, len, s1, s2, so);
    return after_launch("ew2_strided_kernel");
}

template <class F>
int launch_unary(const void* in, void* out, int64_t len, int64_t s1, int64_t so, pncu_stream_t stream) {
    typedef typename F::In In;
    typedef typename F::Out Out;
    if (len <= 0) return 0;
    int rc = ensure_init();
    if (rc) return rc;
    if (!in || !out) { set_error("null operand"); return PNCU_ERR_ARGUMENT; }
    if (((uintptr_t)in | (uint64_t)s1) % sizeof(In) || ((uintptr_t)out | (uint64_t)so) % sizeof(Out)) {
        set_error("operand pointer or stride is not a multiple of the item size");
        return PNCU_ERR_UNSUPPORTED;
    }
    if (len > ((int64_t)1 << 38)) { set_error("length %lld exceeds the grid limit", (long long)len); return PNCU_ERR_ARGUMENT; }
    if ((rc = check_alias(in, s1, sizeof(In), out, so, sizeof(Out), len))) return rc;
    cudaStream_t st = as_stream(stream);

    constexpr int EPV = Epv<In, Out>::value;
    constexpr int IN_VB = EPV * (int)sizeof(In);
    constexpr int OUT_VB = EPV * (int)sizeof(Out);
    if (so == (int64_t)sizeof(Out) && s1 == (int64_t)sizeof(In)) {
        int64_t head = ((OUT_VB - (int64_t)((uintptr_t)out % OUT_VB)) % OUT_VB) / (int64_t)sizeof(Out);
        if (head > len) head = len;
        const unsigned ma = (unsigned)(((uintptr_t)in + head * sizeof(In)) % IN_VB);
        if (ma == 0) {
            const int64_t nvec = (len - head) / EPV;
            if constexpr (smem_of<F>::value != 0) {
                ew1_table_kernel<F><<<full_grid(nvec, EW_BLOCK * unroll_of<F>::value * EW_TABLE_TILES), EW_BLOCK, 0, st>>>(
                    (const In*)in, (Out*)out, head, nvec, len);
                return after_launch("ew1_table_kernel");
            } else {
                ew1_vec_kernel<F><<<full_grid(nvec, EW_BLOCK * unroll_of<F>::value), EW_BLOCK, 0, st>>>((const In*)in, (Out*)out, head, nvec, len);
                return after_launch("ew1_vec_kernel");
            }
        }
        if constexpr (IN_VB % 4 == 0) {
            // input misaligned relative to the output: realigning vector kernel (see ld_pack_realigned); one more
            // vector goes to the scalar head / tail so that no lane reads outside the array
            if ((int64_t)(head * sizeof(In)) < (int64_t)ma) head += EPV;
            if (head > len) head = len;
            int64_t nvec = (len - head) / EPV - 1;
            if (nvec < 0) nvec = 0;
            ew1_realign_kernel<F><<<full_grid(nvec, EW_RL_VPB * unroll_of<F>::value), EW_BLOCK, 0, st>>>((const In*)in, (Out*)out, head, nvec, len, ma);
            return after_launch("ew1_realign_kernel");
        }
    }
    ew1_strided_kernel<F><<<full_grid(len, EW_BLOCK), EW_BLOCK, 0, st>>>((const char*)in, (char*)out, len, s1, so);
    return after_launch("ew1_strided_kernel");
}

}  // namespace pncu
