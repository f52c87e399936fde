// Reductions and arg-reductions: pncu_GetReduceMathOpFast, pncu_GetArgMinMaxOpFast and the
// two-pass building blocks pncu_reduce_partials / pncu_reduce_finish (+ argminmax forms).
//
// replaces: ReduceMathOpFast / ReduceMathOpFastUpcast / GetReduceMathOpFast
// (src/atop/ops_binary.cpp:186-344, 1060-1127), the reduce branch of the dispatcher with its
// per-16K-chunk partials and second pass on the caller (src/pnumpy/_pnumpy.cpp:646-731,
// 359-444) and the arg-reduction pass-through (src/pnumpy/_pnumpy.cpp:1105-1118).
//
// Scheme.  [0, n) is cut into fixed LOGICAL BLOCKS of 64 KiB (65536 / itemsize elements).
//   pass 1: FULL grid, one 256-thread CTA per logical block (the sweep in tools/stream_tune.cu
//           showed a full grid of small CTAs beats a persistent grid by ~10% on B200: the
//           hardware CTA scheduler absorbs the tail).  A thread issues its eight 256-bit loads
//           (256 B in flight per thread) before the first use, folds them into four independent
//           accumulators in a fixed order, then a shuffle-down tree per warp, then the warp
//           leaders through shared memory folded by thread 0 in warp order -> partial[k] in slot k.
//           Arg-reductions keep the block in registers: phase A finds the block extreme
//           (NaN-propagating), phase B scans the registers for its first occurrence; HBM is read
//           once and there is no per-element index arithmetic.
//   pass 2: ONE CTA folds the partials in two levels.  Level A: every GROUP of 64 consecutive slots (4 MiB of
//           input) is folded by eight lanes -- lane j takes slots 64g + j, + 8, ... in order, then a three-level
//           shuffle-down tree -- into group partial g.  Level B: (virtual) thread t of 1024 takes group partials t, t+1024,
//           ... in order, then the same warp/shared tree, then the start value is applied and the result written.
//   ONE LAUNCH: the complete reductions (pncu_GetReduceMathOpFast launchers, pncu_*_fused) do not launch
//           pass 2 at all.  Every pass-1 CTA takes a ticket after storing its partial (one atomicAdd, used
//           for NOTHING but the ticket); the CTA that draws the last ticket runs pass 2 itself, its 256
//           threads standing in for the 1024 of the finish kernel (thread t plays t, t+256, t+512, t+768),
//           so the fold order -- and every bit of the result -- is that of the two-launch building blocks.
//           With several GPUs the last CTA of each first hands its shard to the consumers over NVLink (one system
//           fence, one arrival count per destination) and the consumer's last CTA waits on the device for the other
//           shards before it finishes.  What is handed over: when every shard starts at a multiple of 64 slots,
//           its GROUP partials (level A done by the owner: 1/64 of the bytes, and the consumer only runs level B);
//           otherwise the raw slots (the consumer runs both levels over the whole table).  Same tree either way.
// Every choice above is a function of (n, type) only -- not of the grid, the SM count, the shard boundaries or the
// arrival order of CTAs -- so results are bit-reproducible run to run, and a multi-GPU split
// at multiples of 65536 elements is bit-identical to the single-GPU result.  No atomics anywhere.
//
// Semantics (8c):
//   sum/prod of float32 accumulate in float64 (the reference's intent, ops_binary.cpp:1069-70),
//   rounded to float32 once at the end; float64 in float64; integers wrap modulo 2^k; bool sum
//   is logical OR and bool prod logical AND (ops_binary.cpp:1066,1082).
//   min/max follow NumPy: any NaN -> NaN (the reference has no float body and runs NumPy's
//   loop per chunk, ops_binary.cpp:1096-1098).
//   argmax/argmin follow NumPy: first occurrence of the extreme; for floats the first NaN wins.
#include <float.h>
#include "ew.cuh"
#include "functors.cuh"

using namespace pncu;

namespace {

constexpr int RED_BLOCK_BYTES = 65536;   // one logical block
constexpr int64_t RED_ALIGN_ELEMS = 65536;  // shard boundaries: a multiple of every type's block
constexpr int RED_THREADS = 256;
constexpr int RED_NV = RED_BLOCK_BYTES / (RED_THREADS * 32);  // 256-bit vectors per thread (8)
constexpr int RED_CHAINS = 4;            // independent accumulators per thread
constexpr int FIN_THREADS = 1024;
constexpr int FOLD_GROUP = 64;           // slots per level-A group of pass 2
inline __host__ __device__ int64_t groups_for(int64_t slots) { return (slots + FOLD_GROUP - 1) / FOLD_GROUP; }
template <class In> __host__ __device__ constexpr int64_t block_elems() { return RED_BLOCK_BYTES / (int64_t)sizeof(In); }

// ---- reducer policies ------------------------------------------------------------------------
// In: element type; Acc: accumulator == partial type;
// first(x): accumulator seeded from an element; step(acc, x, idx); combine(a, b) with `a` the
// earlier operand; start(x): accumulator from the caller's start value; result(acc) -> In.

template <class T, class A>
struct SumR {
    typedef T In; typedef A Acc;
    static constexpr bool kHasIdentity = true;
    static __device__ __forceinline__ A identity() { return (A)0; }
    static __device__ __forceinline__ A first(T x, int64_t) { return (A)x; }
    static __device__ __forceinline__ A step(A acc, T x, int64_t) { return acc + (A)x; }
    static __device__ __forceinline__ A combine(A a, A b) { return a + b; }
    static __device__ __forceinline__ A start(T x) { return (A)x; }
    static __device__ __forceinline__ T result(A a) { return (T)a; }
};
template <class T, class A>
struct ProdR {
    typedef T In; typedef A Acc;
    static constexpr bool kHasIdentity = true;
    static __device__ __forceinline__ A identity() { return (A)1; }
    static __device__ __forceinline__ A first(T x, int64_t) { return (A)x; }
    static __device__ __forceinline__ A step(A acc, T x, int64_t) { return acc * (A)x; }
    static __device__ __forceinline__ A combine(A a, A b) { return a * b; }
    static __device__ __forceinline__ A start(T x) { return (A)x; }
    static __device__ __forceinline__ T result(A a) { return (T)a; }
};
struct AnyR {  // bool sum
    typedef bool8 In; typedef unsigned Acc;
    static constexpr bool kHasIdentity = true;
    static __device__ __forceinline__ unsigned identity() { return 0u; }
    static __device__ __forceinline__ unsigned first(bool8 x, int64_t) { return x != 0; }
    static __device__ __forceinline__ unsigned step(unsigned acc, bool8 x, int64_t) { return acc | (unsigned)(x != 0); }
    static __device__ __forceinline__ unsigned combine(unsigned a, unsigned b) { return a | b; }
    static __device__ __forceinline__ unsigned start(bool8 x) { return x != 0; }
    static __device__ __forceinline__ bool8 result(unsigned a) { return (bool8)(a != 0); }
};
struct AllR {  // bool prod
    typedef bool8 In; typedef unsigned Acc;
    static constexpr bool kHasIdentity = true;
    static __device__ __forceinline__ unsigned identity() { return 1u; }
    static __device__ __forceinline__ unsigned first(bool8 x, int64_t) { return x != 0; }
    static __device__ __forceinline__ unsigned step(unsigned acc, bool8 x, int64_t) { return acc & (unsigned)(x != 0); }
    static __device__ __forceinline__ unsigned combine(unsigned a, unsigned b) { return a & b; }
    static __device__ __forceinline__ unsigned start(bool8 x) { return x != 0; }
    static __device__ __forceinline__ bool8 result(unsigned a) { return (bool8)(a != 0); }
};
// sub-word integers are folded in a 32-bit register (byte/short accumulator arrays would
// otherwise be demoted to local memory)
template <class T> struct Widen { typedef T type; };
template <> struct Widen<int8_t> { typedef int type; };
template <> struct Widen<int16_t> { typedef int type; };
template <> struct Widen<uint8_t> { typedef unsigned type; };
template <> struct Widen<uint16_t> { typedef unsigned type; };

template <class T>
struct MinR {
    typedef T In; typedef typename Widen<T>::type Acc;
    static constexpr bool kHasIdentity = false;
    static __device__ __forceinline__ Acc identity() { return Acc(); }
    static __device__ __forceinline__ Acc first(T x, int64_t) { return (Acc)x; }
    static __device__ __forceinline__ Acc step(Acc acc, T x, int64_t) { return op_min<Acc>(acc, (Acc)x); }
    static __device__ __forceinline__ Acc combine(Acc a, Acc b) { return op_min<Acc>(a, b); }
    static __device__ __forceinline__ Acc start(T x) { return (Acc)x; }
    static __device__ __forceinline__ T result(Acc a) { return (T)a; }
};
template <class T>
struct MaxR {
    typedef T In; typedef typename Widen<T>::type Acc;
    static constexpr bool kHasIdentity = false;
    static __device__ __forceinline__ Acc identity() { return Acc(); }
    static __device__ __forceinline__ Acc first(T x, int64_t) { return (Acc)x; }
    static __device__ __forceinline__ Acc step(Acc acc, T x, int64_t) { return op_max<Acc>(acc, (Acc)x); }
    static __device__ __forceinline__ Acc combine(Acc a, Acc b) { return op_max<Acc>(a, b); }
    static __device__ __forceinline__ Acc start(T x) { return (Acc)x; }
    static __device__ __forceinline__ T result(Acc a) { return (T)a; }
};

// (value, index) pair of an arg-reduction.  16 bytes for every T: value at 0, index at 8.
template <class T>
struct alignas(16) ArgPair {
    T v;
    int64_t idx;
};
template <class T, bool IS_MAX>
struct ArgR {
    typedef T In; typedef ArgPair<T> Acc;
    static constexpr bool kHasIdentity = false;
    static constexpr bool kIsMax = IS_MAX;
    static __device__ __forceinline__ Acc identity() { Acc a; a.v = T(); a.idx = 0; return a; }
    static __device__ __forceinline__ Acc first(T x, int64_t idx) { Acc a; a.v = x; a.idx = idx; return a; }
    // strictly better than the incumbent; a NaN beats any non-NaN, nothing beats a NaN
    static __device__ __forceinline__ bool better(T x, T cur) {
        if constexpr (is_fp<T>::value) {
            return (IS_MAX ? x > cur : x < cur) || (x != x && cur == cur);
        } else {
            return IS_MAX ? x > cur : x < cur;
        }
    }
    static __device__ __forceinline__ Acc step(Acc acc, T x, int64_t idx) {
        // a thread visits indices in increasing order, so a strict test keeps the first hit
        if (better(x, acc.v)) { acc.v = x; acc.idx = idx; }
        return acc;
    }
    static __device__ __forceinline__ Acc combine(Acc a, Acc b) {
        if (better(b.v, a.v)) return b;
        if (better(a.v, b.v)) return a;
        return a.idx <= b.idx ? a : b;  // equal (or both NaN): smaller index
    }
};

// ---- shuffles for every accumulator type ---------------------------------------------------------
template <class A> __device__ __forceinline__ A shfl_down_any(A a, int o) { return __shfl_down_sync(0xffffffffu, a, o); }
template <> __device__ __forceinline__ int8_t shfl_down_any(int8_t a, int o) { return (int8_t)__shfl_down_sync(0xffffffffu, (int)a, o); }
template <> __device__ __forceinline__ uint8_t shfl_down_any(uint8_t a, int o) { return (uint8_t)__shfl_down_sync(0xffffffffu, (unsigned)a, o); }
template <> __device__ __forceinline__ int16_t shfl_down_any(int16_t a, int o) { return (int16_t)__shfl_down_sync(0xffffffffu, (int)a, o); }
template <> __device__ __forceinline__ uint16_t shfl_down_any(uint16_t a, int o) { return (uint16_t)__shfl_down_sync(0xffffffffu, (unsigned)a, o); }
template <> __device__ __forceinline__ int64_t shfl_down_any(int64_t a, int o) { return (int64_t)__shfl_down_sync(0xffffffffu, (long long)a, o); }
template <> __device__ __forceinline__ uint64_t shfl_down_any(uint64_t a, int o) { return (uint64_t)__shfl_down_sync(0xffffffffu, (unsigned long long)a, o); }
template <class T> __device__ __forceinline__ ArgPair<T> shfl_down_pair(ArgPair<T> a, int o) {
    ArgPair<T> r;
    r.v = shfl_down_any<T>(a.v, o);
    r.idx = shfl_down_any<int64_t>(a.idx, o);
    return r;
}
template <class A> struct Shfl { static __device__ __forceinline__ A down(A a, int o) { return shfl_down_any<A>(a, o); } };
template <class T> struct Shfl<ArgPair<T>> { static __device__ __forceinline__ ArgPair<T> down(ArgPair<T> a, int o) { return shfl_down_pair<T>(a, o); } };

// fold one accumulator per thread across the CTA; result valid in thread 0
template <class R, int THREADS>
__device__ __forceinline__ typename R::Acc block_fold(typename R::Acc a, typename R::Acc* smem) {
    typedef typename R::Acc Acc;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a = R::combine(a, Shfl<Acc>::down(a, o));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) smem[warp] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        a = smem[0];
        for (int w = 1; w < THREADS / 32; ++w) a = R::combine(a, smem[w]);
    }
    return a;
}

template <class R>
__device__ __forceinline__ void write_result(typename R::Acc total, typename R::In* out, const typename R::In* startVal) {
    typename R::Acc t = total;
    if (startVal) t = R::combine(R::start(*startVal), total);
    *out = R::result(t);
}

// A block partial is written once per 64 KiB of input and read back by the tail after the whole stream has passed
// through L2, i.e. from HBM (~1 us per dependent round trip on the one SM that runs the tail).  Stored with the L2
// evict_last policy the slots (<= 1 MiB) survive the stream in the 126 MB L2 and the tail's round trips are L2's
// (measured: tails 1-3 us shorter, e.g. float32 sum 6.8 -> 5.4 us, float32 argmax 16.5 -> 13.2 us at 2^28 elements).
template <class A> __device__ __forceinline__ void st_partial(A* p, const A& v) {
    unsigned long long policy;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(policy));
    if constexpr (sizeof(A) == 4) {
        unsigned w;
        memcpy(&w, &v, 4);
        asm volatile("st.global.L2::cache_hint.b32 [%0], %1, %2;" :: "l"(p), "r"(w), "l"(policy) : "memory");
    } else if constexpr (sizeof(A) == 8) {
        unsigned long long w;
        memcpy(&w, &v, 8);
        asm volatile("st.global.L2::cache_hint.b64 [%0], %1, %2;" :: "l"(p), "l"(w), "l"(policy) : "memory");
    } else {
        static_assert(sizeof(A) == 16, "partials are 4, 8 or 16 bytes");
        unsigned long long w[2];
        memcpy(w, &v, 16);
        asm volatile("st.global.L2::cache_hint.v2.b64 [%0], {%1,%2}, %3;" :: "l"(p), "l"(w[0]), "l"(w[1]), "l"(policy) : "memory");
    }
}

// ---- pass 2 as a device function -------------------------------------------------------------------------
// loads that bypass L1: the table was written by other SMs (or other GPUs) during this kernel
template <class A> __device__ __forceinline__ A ld_cg(const A* p) {
    A r;
    if constexpr (sizeof(A) == 4) {
        unsigned v;
        asm volatile("ld.global.cg.b32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
        memcpy(&r, &v, 4);
    } else if constexpr (sizeof(A) == 8) {
        unsigned long long v;
        asm volatile("ld.global.cg.b64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
        memcpy(&r, &v, 8);
    } else {
        static_assert(sizeof(A) == 16, "partials are 4, 8 or 16 bytes");
        unsigned long long v[2];
        asm volatile("ld.global.cg.v2.b64 {%0,%1}, [%2];" : "=l"(v[0]), "=l"(v[1]) : "l"(p) : "memory");
        memcpy(&r, v, 16);
    }
    return r;
}

// Fold `count` partials in the order of a 1024-thread finish CTA, with THREADS real threads: real thread t
// plays the virtual threads t + q*THREADS.  Virtual thread v takes slots v, v+1024, ... in ascending order;
// threads beyond `count` hold a copy of slot 0 (harmless for idempotent ops) or the identity; then a
// shuffle-down tree per (virtual) warp, the 32 warp values through shared memory, folded by thread 0 in warp
// order.  Result valid in thread 0.  `smem`: 32 entries.
template <class R, int THREADS>
__device__ __forceinline__ typename R::Acc fold_table(const typename R::Acc* partials, int64_t count, typename R::Acc* smem) {
    typedef typename R::Acc Acc;
    constexpr int VT = FIN_THREADS / THREADS;
    constexpr int MLP = VT == 1 ? 8 : (sizeof(Acc) >= 16 ? 2 : 4);   // loads in flight per virtual thread
    Acc a[VT];
#pragma unroll
    for (int q = 0; q < VT; ++q) {
        const int64_t v = (int64_t)threadIdx.x + q * THREADS;
        if (R::kHasIdentity) a[q] = R::identity();
        else a[q] = ld_cg(partials + (v < count ? v : 0));   // the first slot replaces the seed
    }
    int64_t k = (int64_t)threadIdx.x + (R::kHasIdentity ? 0 : FIN_THREADS);
    // full batches: unconditional loads first (predicated ones get interleaved with the adds and serialise)
    for (; k + (int64_t)(VT - 1) * THREADS + (int64_t)(MLP - 1) * FIN_THREADS < count; k += (int64_t)FIN_THREADS * MLP) {
        Acc x[VT][MLP];
#pragma unroll
        for (int q = 0; q < VT; ++q)
#pragma unroll
            for (int m = 0; m < MLP; ++m) x[q][m] = ld_cg(partials + k + q * THREADS + (int64_t)m * FIN_THREADS);
#pragma unroll
        for (int q = 0; q < VT; ++q)
#pragma unroll
            for (int m = 0; m < MLP; ++m) a[q] = R::combine(a[q], x[q][m]);
    }
#pragma unroll
    for (int q = 0; q < VT; ++q)
        for (int64_t kk = k + q * THREADS; kk < count; kk += FIN_THREADS) a[q] = R::combine(a[q], ld_cg(partials + kk));
#pragma unroll
    for (int q = 0; q < VT; ++q) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a[q] = R::combine(a[q], Shfl<Acc>::down(a[q], o));
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
#pragma unroll
        for (int q = 0; q < VT; ++q) smem[warp + q * (THREADS / 32)] = a[q];
    }
    __syncthreads();
    Acc r = a[0];
    if (threadIdx.x == 0) {
        r = smem[0];
        for (int w = 1; w < FIN_THREADS / 32; ++w) r = R::combine(r, smem[w]);
    }
    return r;
}

// Level A of pass 2: group g = slots [64g, 64g + 64) of `slots[0, count)` -> store(g, partial).  EIGHT LANES fold one
// group: lane j of the eight takes slots 64g + j, + 8, ... + 56 in ascending order (seven combines in registers),
// then a three-level shuffle-down tree over the eight lanes.  A warp therefore folds four groups per step with three
// shuffle levels in all -- pass 2 runs on ONE SM, whose shuffle unit is the bottleneck of a full-warp tree per group
// (measured: 2x the time of the flat fold it replaced) -- and every load instruction still reads four runs of
// 8 consecutive slots (whole 32-byte sectors for every partial size).  Slots past `count` (last group only) are
// skipped; a lane with none holds the identity, or a copy of the group's first slot for the idempotent ops that have
// none.  STEPS steps are loaded before the first combine (8 STEPS loads in flight per lane); any number of warps may
// share the work (a group's tree does not depend on who runs it).  The destination may be peer memory.
constexpr int FOLD_LANES = 8;                        // lanes per group
constexpr int FOLD_SEQ = FOLD_GROUP / FOLD_LANES;    // slots a lane folds in registers
template <class R, int THREADS, class Store>
__device__ __forceinline__ void fold_groups(const typename R::Acc* slots, int64_t count, Store store) {
    typedef typename R::Acc Acc;
    constexpr int NW = THREADS / 32;
    constexpr int GPW = 32 / FOLD_LANES;             // groups per warp step (4)
    constexpr int STEPS = sizeof(Acc) >= 16 ? 1 : (sizeof(Acc) == 8 ? 2 : 4);   // 32 registers of loads in flight per lane
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane / FOLD_LANES, j = lane % FOLD_LANES;
    const int64_t ngroups = groups_for(count);
    const int64_t full = count / FOLD_GROUP;          // groups with all 64 slots
    for (int64_t g0 = (int64_t)warp * (GPW * STEPS); g0 < ngroups; g0 += (int64_t)NW * (GPW * STEPS)) {
        if (g0 + GPW * STEPS <= full) {
            // unconditional loads first, all of them (predicated loads get interleaved with their uses and serialise;
            // the slots were evicted from L2 by the stream, so every round trip is HBM's)
            Acc x[STEPS][FOLD_SEQ];
#pragma unroll
            for (int st = 0; st < STEPS; ++st)
#pragma unroll
                for (int k = 0; k < FOLD_SEQ; ++k) x[st][k] = ld_cg(slots + (g0 + st * GPW + sub) * FOLD_GROUP + k * FOLD_LANES + j);
#pragma unroll
            for (int st = 0; st < STEPS; ++st) {
                Acc a = x[st][0];
#pragma unroll
                for (int k = 1; k < FOLD_SEQ; ++k) a = R::combine(a, x[st][k]);
#pragma unroll
                for (int o = FOLD_LANES / 2; o > 0; o >>= 1) a = R::combine(a, Shfl<Acc>::down(a, o));
                if (j == 0) store(g0 + st * GPW + sub, a);
            }
        } else {
            // the last batch of a warp: step by step, the ragged last group with its fill-ins
            for (int st = 0; st < STEPS; ++st) {
                const int64_t g = g0 + st * GPW + sub;
                const int64_t base = g * FOLD_GROUP;
                const bool live = base < count;          // (all lanes run the shuffles)
                Acc a = R::identity();
                if (live) {
                    const Acc fill = R::kHasIdentity ? R::identity() : ld_cg(slots + base);
                    a = base + j < count ? ld_cg(slots + base + j) : fill;
                    for (int k = 1; k < FOLD_SEQ; ++k) {
                        const int64_t i = base + k * FOLD_LANES + j;
                        if (i < count) a = R::combine(a, ld_cg(slots + i));
                    }
                }
#pragma unroll
                for (int o = FOLD_LANES / 2; o > 0; o >>= 1) a = R::combine(a, Shfl<Acc>::down(a, o));
                if (live && j == 0) store(g, a);
            }
        }
    }
}

// Only floating-point sums and products depend on the fold ORDER (integers wrap, min / max / arg-reductions are
// associative and commutative bit for bit), so only they must take the two-level tree wherever the table is folded.
// Everything else folds a whole table FLAT when one CTA has it all in front of it anyway (one level less: ~2-4 us on a
// 2^28-element reduction) and in two levels when that spreads the work over the GPUs -- same bits by associativity.
template <class R> struct OrderSensitive { static constexpr bool value = false; };
template <class T> struct OrderSensitive<SumR<T, double>> { static constexpr bool value = true; };
template <class T> struct OrderSensitive<ProdR<T, double>> { static constexpr bool value = true; };

// both levels by one CTA: level A into `gscratch` (global memory of this GPU, groups_for(count) entries), level B
// over it.  Result valid in thread 0.
template <class R, int THREADS>
__device__ __forceinline__ typename R::Acc fold_two_level(const typename R::Acc* slots, int64_t count, typename R::Acc* gscratch,
                                                          typename R::Acc* smem) {
    if constexpr (!OrderSensitive<R>::value) return fold_table<R, THREADS>(slots, count, smem);
    fold_groups<R, THREADS>(slots, count, [gscratch](int64_t g, typename R::Acc a) { gscratch[g] = a; });
    __syncthreads();   // the CTA's own global stores are visible to all of its threads (read back through L2)
    return fold_table<R, THREADS>(gscratch, groups_for(count), smem);
}

template <class R, bool VALUE>
__device__ __forceinline__ void store_result(typename R::Acc a, void* out, const void* startVal) {
    if constexpr (VALUE) {
        write_result<R>(a, reinterpret_cast<typename R::In*>(out), reinterpret_cast<const typename R::In*>(startVal));
    } else {
        if constexpr (sizeof(typename R::Acc) == 16) *reinterpret_cast<int64_t*>(out) = a.idx;
    }
}

// ---- the tail of a ONE-LAUNCH reduction ------------------------------------------------------------------
// What the last CTA of a pass-1 grid needs to finish the reduction (and, with n > 1 participants, to exchange
// the shard's partials first).  Passed by value.
struct FusedArgs {
    unsigned* ticket;                 // this GPU's ticket counter: 0 between launches (the last CTA re-arms it)
    int n, self, all, root;           // participants; `all`: every participant folds (all-reduce), else only `root`
    long long slot_base, total_count; // this shard's first slot in a consumer's table; slots of all shards
    void* slots[PNCU_MAX_PEERS];      // consumers' slot tables (peer memory); slots[self] + slot_base is where a
                                      // consuming participant's own pass 1 writes directly
    unsigned long long* arrived[PNCU_MAX_PEERS];
    unsigned long long expected;      // arrivals a consumer's counter shows once every other shard is in
    long long group_off;              // > 0: hierarchical -- every shard starts at a multiple of FOLD_GROUP slots and
                                      // consumer d's GROUP table (groups_for(total_count) entries) is slots[d] + group_off bytes
    void* gscratch;                   // this GPU's level-A output (group partials of the slots it folds itself)
    unsigned* gtickets;               // != NULL: GROUP TICKETS -- one counter per group of FOLD_GROUP slots of THIS launch, zero
                                      // between launches; the CTA that completes a group folds it during pass 1 (see fused_tail)
    void* out;
    const void* startVal;
    unsigned long long* done_flag;    // optional (may be mapped host memory): receives done_value after `out`,
    unsigned long long done_value;    //   or done_value | PNCU_FUSED_TIMEOUT_BIT when a shard never arrived
};
__device__ unsigned long long g_exchange_timeouts = 0;
// duration of the most recent one-launch tail on this device (last CTA: ticket drawn -> result published), for
// tools/reduce_breakdown.py: two %globaltimer reads and one store per KERNEL
__device__ unsigned long long g_tail_ns = 0;
__device__ __forceinline__ unsigned long long global_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
constexpr unsigned long long EXCH_WAIT_NS = 2000000000ULL;

// thread 0: wait until *counter >= expected; false after EXCH_WAIT_NS (a lost participant is reported, not hung on)
__device__ __forceinline__ bool wait_arrivals(const unsigned long long* counter, unsigned long long expected) {
    unsigned long long t0, t1, v;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(counter) : "memory");
        if (v >= expected) return true;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > EXCH_WAIT_NS) { atomicAdd(&g_exchange_timeouts, 1ULL); return false; }
        __nanosleep(64);
    }
}
__device__ __forceinline__ void publish(unsigned long long* flag, unsigned long long value) {
    if (flag) asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(flag), "l"(value) : "memory");
}

// The tail runs ONCE per kernel, on one CTA, with its code cold (the stream has evicted it from every cache): its
// instructions are fetched from HBM, so its size is part of its latency (measured: a tail with three inlined copies of
// each level cost ~7 us more than the same arithmetic in one copy).  Hence: one out-of-line instance of level A
// (destinations passed as a list) and one of level B, shared by the single-GPU, hierarchical and flat paths.
// where level A's group partials go: this GPU's scratch, or (hierarchical exchange) every consumer's group table
template <class R>
__device__ __forceinline__ void fused_level_a(const typename R::Acc* slots, int64_t count, const FusedArgs& fa, bool hier) {
    typedef typename R::Acc Acc;
    const int64_t gbase = fa.slot_base / FOLD_GROUP;
    fold_groups<R, RED_THREADS>(slots, count, [&fa, hier, gbase](int64_t g, Acc a) {
        if (!hier) { reinterpret_cast<Acc*>(fa.gscratch)[g] = a; return; }
        for (int d = 0; d < fa.n; ++d)
            if (fa.all || d == fa.root) reinterpret_cast<Acc*>(reinterpret_cast<char*>(fa.slots[d]) + fa.group_off)[gbase + g] = a;
    });
}
template <class R>
__device__ __forceinline__ typename R::Acc fused_level_b(const typename R::Acc* groups, int64_t ngroups, typename R::Acc* smem) {
    return fold_table<R, RED_THREADS>(groups, ngroups, smem);
}

template <class R, bool VALUE>
__device__ __noinline__ void fused_finish(const typename R::Acc* local, int64_t local_count, const FusedArgs& fa) {
    typedef typename R::Acc Acc;
    __shared__ Acc fin_smem[FIN_THREADS / 32];
    __shared__ int s_ok;
    const unsigned long long t_begin = global_ns();
    __threadfence();                         // the other CTAs' partials (acquire side of their tickets)
    if (threadIdx.x == 0) *fa.ticket = 0u;   // re-arm for the next launch on this stream
    const bool multi = fa.n > 1;
    const bool consumer = !multi || fa.all || fa.self == fa.root;
    const bool hier = multi && fa.group_off > 0;
    const bool groups_done = fa.gtickets != nullptr;   // level A of the local slots happened during pass 1 (fused_tail)
    // does this CTA fold a whole slot table by itself (single GPU, or the consumer of a flat exchange)?  Then only the
    // order-sensitive reductions need level A; the others fold the slots flat (see OrderSensitive)
    const bool two_level = hier || groups_done || OrderSensitive<R>::value;
    const Acc* table = local;                // what level A runs over on this GPU, and how many slots
    int64_t count = local_count;
    if (groups_done) {
        if (hier) {   // hand the local group partials to every consumer's group table (NVLink)
            const int64_t gbase = fa.slot_base / FOLD_GROUP, ngl = groups_for(local_count);
            const Acc* gl = reinterpret_cast<const Acc*>(fa.gscratch);
            for (int64_t i = threadIdx.x; i < ngl; i += RED_THREADS) {
                const Acc v = ld_cg(gl + i);
                for (int d = 0; d < fa.n; ++d)
                    if (fa.all || d == fa.root) reinterpret_cast<Acc*>(reinterpret_cast<char*>(fa.slots[d]) + fa.group_off)[gbase + i] = v;
            }
        }
    } else if (multi && !hier) {
        // flat: push the raw slots, PUSH_MLP loads in flight per thread, then the stores to every consumer
        constexpr int PUSH_MLP = 8;
        for (int64_t i0 = threadIdx.x; i0 < local_count; i0 += (int64_t)RED_THREADS * PUSH_MLP) {
            Acc x[PUSH_MLP];
#pragma unroll
            for (int m = 0; m < PUSH_MLP; ++m) {
                const int64_t i = i0 + (int64_t)m * RED_THREADS;
                if (i < local_count) x[m] = ld_cg(local + i);
            }
            for (int d = 0; d < fa.n; ++d) {
                if (d == fa.self || !(fa.all || d == fa.root)) continue;
                Acc* to = reinterpret_cast<Acc*>(fa.slots[d]) + fa.slot_base;
#pragma unroll
                for (int m = 0; m < PUSH_MLP; ++m) {
                    const int64_t i = i0 + (int64_t)m * RED_THREADS;
                    if (i < local_count) to[i] = x[m];
                }
            }
        }
        table = reinterpret_cast<const Acc*>(fa.slots[fa.self]);   // a consumer runs level A over the WHOLE table later
        count = fa.total_count;
    } else if (two_level) {
        fused_level_a<R>(table, count, fa, hier);
    }
    if (multi) {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence_system();
            for (int d = 0; d < fa.n; ++d)
                if (d != fa.self && (fa.all || d == fa.root)) atomicAdd_system(fa.arrived[d], 1ULL);
            s_ok = 1;
            if (consumer) s_ok = wait_arrivals(fa.arrived[fa.self], fa.expected) ? 1 : 0;
        }
        __syncthreads();
        if (!consumer) return;
        if (!s_ok) {   // never fold an incomplete table: report, leave `out` alone
            if (threadIdx.x == 0) publish(fa.done_flag, fa.done_value | PNCU_FUSED_TIMEOUT_BIT);
            return;
        }
        if (!hier && two_level) fused_level_a<R>(table, count, fa, hier);
    }
    __syncthreads();   // level A's stores (this CTA's own, or the peers' behind the arrival counter) before level B's loads
    const Acc* groups = hier ? reinterpret_cast<const Acc*>(reinterpret_cast<const char*>(fa.slots[fa.self]) + fa.group_off)
                             : reinterpret_cast<const Acc*>(fa.gscratch);
    const Acc a = two_level ? fused_level_b<R>(groups, groups_for(multi ? fa.total_count : local_count), fin_smem)
                            : fused_level_b<R>(table, count, fin_smem);
    if (threadIdx.x == 0) {
        store_result<R, VALUE>(a, fa.out, fa.startVal);
        publish(fa.done_flag, fa.done_value);
        g_tail_ns = global_ns() - t_begin;
    }
}

// One group of pass 2's level A, folded by the first eight lanes of ONE warp (call with a whole warp): exactly the tree
// of fold_groups -- lane j takes slots j, j + 8, ... in order, then the three-level shuffle tree.  Result in lane 0.
template <class R>
__device__ __forceinline__ typename R::Acc fold_group_warp(const typename R::Acc* gslots, int cnt) {
    typedef typename R::Acc Acc;
    const int j = threadIdx.x & 31;
    Acc a = R::identity();
    if (j < FOLD_LANES) {
        if (cnt == FOLD_GROUP) {
            Acc x[FOLD_SEQ];
#pragma unroll
            for (int k = 0; k < FOLD_SEQ; ++k) x[k] = ld_cg(gslots + k * FOLD_LANES + j);
            a = x[0];
#pragma unroll
            for (int k = 1; k < FOLD_SEQ; ++k) a = R::combine(a, x[k]);
        } else {
            const Acc fill = R::kHasIdentity ? R::identity() : ld_cg(gslots);
            a = j < cnt ? ld_cg(gslots + j) : fill;
            for (int k = 1; k < FOLD_SEQ; ++k) {
                const int i = k * FOLD_LANES + j;
                if (i < cnt) a = R::combine(a, ld_cg(gslots + i));
            }
        }
    }
#pragma unroll
    for (int o = FOLD_LANES / 2; o > 0; o >>= 1) a = R::combine(a, Shfl<Acc>::down(a, o));
    return a;
}

// called by every CTA of a FUSED pass-1 grid after thread 0 stored partials[blockIdx.x].
// Classic form (fa.gtickets == NULL): one ticket per CTA; the CTA that draws the last one runs all of pass 2.
// GROUP TICKETS: a CTA draws a ticket of ITS GROUP of 64 slots; the one that completes the group folds it right away
// (one warp, eight loads in flight, the slots still hot in L2) into gscratch[g] and then draws a ticket of the grid;
// whoever completes the grid finds level A done and only runs level B -- over 1/64 of the entries.  The tail of a
// 2^28-element float32 sum shrinks from ~5 us to ~2 us, that of an arg-reduction (16-byte partials) from ~15 us; the
// price is ~0.6 us in one CTA out of 64.  Same tree, same bits.
template <class R, bool VALUE>
__device__ __forceinline__ void fused_tail(const typename R::Acc* partials, const FusedArgs& fa) {
    typedef typename R::Acc Acc;
    __shared__ int s_role;   // 0: nothing left to do; 1: fold my group; 2: finish the reduction
    __shared__ int s_last;   // (after a group fold) this CTA completed the grid
    const unsigned g = blockIdx.x / FOLD_GROUP;
    const unsigned in_group = min((unsigned)FOLD_GROUP, gridDim.x - g * FOLD_GROUP);
    if (threadIdx.x == 0) {
        __threadfence();
        if (fa.gtickets) s_role = atomicAdd(fa.gtickets + g, 1u) == in_group - 1 ? 1 : 0;
        else s_role = atomicAdd(fa.ticket, 1u) == gridDim.x - 1 ? 2 : 0;
    }
    __syncthreads();
    const int role = s_role;   // (never written again)
    if (role == 0) return;
    if (role == 1) {
        if (threadIdx.x < 32) {
            __threadfence();   // the other CTAs' partials of this group (acquire side of their tickets)
            const Acc a = fold_group_warp<R>(partials + (size_t)g * FOLD_GROUP, (int)in_group);
            if (threadIdx.x == 0) {
                st_partial(reinterpret_cast<Acc*>(fa.gscratch) + g, a);
                fa.gtickets[g] = 0u;   // re-arm
                __threadfence();
                s_last = atomicAdd(fa.ticket, 1u) == (gridDim.x + FOLD_GROUP - 1) / FOLD_GROUP - 1 ? 1 : 0;
            }
        }
        __syncthreads();
        if (!s_last) return;
    }
    fused_finish<R, VALUE>(partials, (int64_t)gridDim.x, fa);
}

// ---- word-at-a-time folds for the 1- and 2-byte types ----------------------------------------------
// Extracting bytes one by one costs ~3 ALU instructions per element.  These fold a whole 32-bit word
// per step with the integer dot-product unit (IDP.4A / IDP.2A: sum of the 4 or 2 packed lanes in one
// instruction) and the packed 16x2 min/max (VIMNMX.S16x2 / .U16x2; bytes are widened to 16x2 by one
// PRMT per pair).  Integer results are exact in any order, so this changes no value.  (Measured: it
// does not change the 2^28-element timings either -- a 256 MiB int8 array takes 38 us to stream and
// the two launches + the second pass add ~25 us; at equal BYTES the small types match the wide ones.)
template <class R> struct WordOps { static constexpr bool enabled = false; };

template <> struct WordOps<SumR<int8_t, unsigned>> {
    static constexpr bool enabled = true; typedef unsigned W;
    static __device__ __forceinline__ W seed(int8_t) { return 0u; }
    static __device__ __forceinline__ W step(W a, uint32_t w) { return (unsigned)__dp4a((int)w, 0x01010101, (int)a); }
    static __device__ __forceinline__ unsigned finish(W a) { return a; }
};
template <> struct WordOps<SumR<uint8_t, unsigned>> {
    static constexpr bool enabled = true; typedef unsigned W;
    static __device__ __forceinline__ W seed(uint8_t) { return 0u; }
    static __device__ __forceinline__ W step(W a, uint32_t w) { return __dp4a((unsigned)w, 0x01010101u, a); }
    static __device__ __forceinline__ unsigned finish(W a) { return a; }
};
template <> struct WordOps<SumR<int16_t, unsigned>> {
    static constexpr bool enabled = true; typedef unsigned W;
    static __device__ __forceinline__ W seed(int16_t) { return 0u; }
    static __device__ __forceinline__ W step(W a, uint32_t w) { return (unsigned)__dp2a_lo((int)w, 0x00000101, (int)a); }
    static __device__ __forceinline__ unsigned finish(W a) { return a; }
};
template <> struct WordOps<SumR<uint16_t, unsigned>> {
    static constexpr bool enabled = true; typedef unsigned W;
    static __device__ __forceinline__ W seed(uint16_t) { return 0u; }
    static __device__ __forceinline__ W step(W a, uint32_t w) { return __dp2a_lo((unsigned)w, 0x00000101u, a); }
    static __device__ __forceinline__ unsigned finish(W a) { return a; }
};
template <> struct WordOps<AnyR> {
    static constexpr bool enabled = true; typedef unsigned W;
    static __device__ __forceinline__ W seed(bool8) { return 0u; }
    static __device__ __forceinline__ W step(W a, uint32_t w) { return a | w; }
    static __device__ __forceinline__ unsigned finish(W a) { return a != 0u; }
};
template <> struct WordOps<AllR> {
    static constexpr bool enabled = true; typedef unsigned W;
    static __device__ __forceinline__ W seed(bool8) { return 0xffffffffu; }
    static __device__ __forceinline__ W step(W a, uint32_t w) { return a & __vcmpne4(w, 0u); }  // 0xff per non-zero byte
    static __device__ __forceinline__ unsigned finish(W a) { return a == 0xffffffffu; }
};
// packed 16x2 min / max
template <bool IS_MAX, bool SIGNED> __device__ __forceinline__ unsigned mm16x2(unsigned a, unsigned b) {
    if (SIGNED) return IS_MAX ? __vmaxs2(a, b) : __vmins2(a, b);
    return IS_MAX ? __vmaxu2(a, b) : __vminu2(a, b);
}
template <class T, bool IS_MAX> struct MinMaxWord {
    static constexpr bool enabled = true; typedef unsigned W;
    static constexpr bool kSigned = std::is_signed<T>::value;
    typedef typename Widen<T>::type A;
    static __device__ __forceinline__ W seed(T first) { return ((unsigned)(uint16_t)(int16_t)(A)first) * 0x00010001u; }
    static __device__ __forceinline__ W step(W a, uint32_t w) {
        if constexpr (sizeof(T) == 2) {
            return mm16x2<IS_MAX, kSigned>(a, w);
        } else {  // bytes -> two 16x2 words (sign- or zero-extended) by PRMT
            // NB: the __byte_perm intrinsic masks the selector to 3 bits per nibble and so cannot ask for
            // sign replication (selector bit 3); the PTX instruction can.
            unsigned lo, hi;
            if (kSigned) {
                asm("prmt.b32 %0, %1, %2, 0x9180;" : "=r"(lo) : "r"(w), "r"(0u));   // {b0, sign b0, b1, sign b1}
                asm("prmt.b32 %0, %1, %2, 0xb3a2;" : "=r"(hi) : "r"(w), "r"(0u));   // {b2, sign b2, b3, sign b3}
            } else {
                lo = __byte_perm(w, 0u, 0x4140);
                hi = __byte_perm(w, 0u, 0x4342);
            }
            return mm16x2<IS_MAX, kSigned>(mm16x2<IS_MAX, kSigned>(a, lo), hi);
        }
    }
    static __device__ __forceinline__ A finish(W a) {
        const A x = kSigned ? (A)(int16_t)(a & 0xffffu) : (A)(a & 0xffffu);
        const A y = kSigned ? (A)(int16_t)(a >> 16) : (A)(a >> 16);
        return IS_MAX ? (x > y ? x : y) : (x < y ? x : y);
    }
};
template <> struct WordOps<MinR<int8_t>> : MinMaxWord<int8_t, false> {};
template <> struct WordOps<MinR<uint8_t>> : MinMaxWord<uint8_t, false> {};
template <> struct WordOps<MinR<int16_t>> : MinMaxWord<int16_t, false> {};
template <> struct WordOps<MinR<uint16_t>> : MinMaxWord<uint16_t, false> {};
template <> struct WordOps<MaxR<int8_t>> : MinMaxWord<int8_t, true> {};
template <> struct WordOps<MaxR<uint8_t>> : MinMaxWord<uint8_t, true> {};
template <> struct WordOps<MaxR<int16_t>> : MinMaxWord<int16_t, true> {};
template <> struct WordOps<MaxR<uint16_t>> : MinMaxWord<uint16_t, true> {};

// ---- a thread's share of one FULL logical block, for any element-aligned block start -----------------------
// Aligned start: pack j of thread t is vector j*256 + t of the block, one 256-bit load each.
// Misaligned start -- NumPy reduces a.min() / a.max() over a[1:], and views like a[1:].sum() -- : the block is
// re-cut at the next 32-byte boundary: `head` = elements before it (0 < head < EPV); the NV-1 aligned vectors
// from there are dealt out exactly as before (still one 256-bit load each, fully coalesced; the very last slot,
// pack 7 of thread 255, stays empty) and the EPV elements left over at the two ends of the block -- the `head`
// leading ones and the EPV - head trailing ones -- are picked up one each by threads 0 .. EPV-1 with an element
// load issued alongside the vector loads.  Same bytes from HBM, same instruction mix.  The fold order is a
// function of (n, type, alignment) only, so determinism and shard-invariance are untouched.
template <class In, int EPV>
__device__ __forceinline__ int block_head(const void* block_start) {
    return (int)(((32u - (unsigned)((uintptr_t)block_start & 31u)) & 31u) / (unsigned)sizeof(In));
}
// position (element offset in the block) of leftover element e
template <int EPV>
__device__ __forceinline__ int leftover_pos(int e, int head) { return e < head ? e : (RED_NV * RED_THREADS - 1) * EPV + e; }

template <class In, int EPV, bool SHIFTED>
__device__ __forceinline__ void load_block(const In* p, int head, Pack<In, EPV> (&r)[RED_NV], In& extra) {
    if constexpr (!SHIFTED) {
#pragma unroll
        for (int j = 0; j < RED_NV; ++j) r[j] = ld_pack<In, EPV>(p + ((int64_t)j * RED_THREADS + threadIdx.x) * EPV);
    } else {
        const In* pa = p + head;
#pragma unroll
        for (int j = 0; j < RED_NV - 1; ++j) r[j] = ld_pack<In, EPV>(pa + ((int64_t)j * RED_THREADS + threadIdx.x) * EPV);
        // the last slot does not exist for thread 255: it re-reads its previous vector so that every lane issues
        // the same instruction; the fold skips it (see last_slot_empty)
        const int jl = threadIdx.x == RED_THREADS - 1 ? RED_NV - 2 : RED_NV - 1;
        r[RED_NV - 1] = ld_pack<In, EPV>(pa + ((int64_t)jl * RED_THREADS + threadIdx.x) * EPV);
        if (threadIdx.x < EPV) extra = p[leftover_pos<EPV>((int)threadIdx.x, head)];
    }
}
// how a pass-1 kernel reads a full block: element loads (strided input), aligned vectors, or the re-cut above
enum { RD_ELEMS = 0, RD_VEC = 1, RD_SHIFTED = 2 };
inline int read_mode(const void* in, bool contiguous) {
    if (!contiguous) return RD_ELEMS;
    return ((uintptr_t)in % 32) == 0 ? RD_VEC : RD_SHIFTED;
}

// ---- pass 1 --------------------------------------------------------------------------------------
#ifndef PNCU_RED_MINB
#define PNCU_RED_MINB 4
#endif
template <class R, int VEC, bool FUSED>
__global__ void __launch_bounds__(RED_THREADS, PNCU_RED_MINB)
reduce_blocks_kernel(const char* in, int64_t n, int64_t stride, int64_t index_base,
                     typename R::Acc* partials, const __grid_constant__ FusedArgs fa) {
    typedef typename R::In In;
    typedef typename R::Acc Acc;
    constexpr int64_t BE = block_elems<In>();
    constexpr int EPV = 32 / (int)sizeof(In);
    __shared__ Acc smem[RED_THREADS / 32];

    const int64_t k = blockIdx.x;
    const int64_t start = k * BE;
    const int64_t cnt = (n - start) < BE ? (n - start) : BE;
    const char* base = in + start * stride;
    const int64_t gidx0 = start + index_base;

    Acc acc[RED_CHAINS];
    if (R::kHasIdentity) {
#pragma unroll
        for (int c = 0; c < RED_CHAINS; ++c) acc[c] = R::identity();
    } else {
        // idempotent ops (min/max): seed every accumulator with the block's first element
        const Acc seed = R::first(*reinterpret_cast<const In*>(base), gidx0);
#pragma unroll
        for (int c = 0; c < RED_CHAINS; ++c) acc[c] = seed;
    }

    if (VEC && cnt == BE) {
        constexpr bool SHIFTED = VEC == RD_SHIFTED;
        const In* p = reinterpret_cast<const In*>(base);
        Pack<In, EPV> r[RED_NV];
        In extra = In();
        const int head = SHIFTED ? block_head<In, EPV>(p) : 0;
        load_block<In, EPV, SHIFTED>(p, head, r, extra);
        const bool last_slot_empty = SHIFTED && threadIdx.x == RED_THREADS - 1;
        if constexpr (WordOps<R>::enabled) {
            // 1- and 2-byte types: fold whole 32-bit words (4 or 2 elements per instruction)
            typename WordOps<R>::W wacc[RED_CHAINS];
#pragma unroll
            for (int c = 0; c < RED_CHAINS; ++c) wacc[c] = WordOps<R>::seed(*reinterpret_cast<const In*>(base));
#pragma unroll
            for (int j = 0; j < RED_NV; ++j) {
                if (SHIFTED && j == RED_NV - 1 && last_slot_empty) continue;
#pragma unroll
                for (int wi = 0; wi < 8; ++wi) wacc[j % RED_CHAINS] = WordOps<R>::step(wacc[j % RED_CHAINS], r[j].w[wi]);
            }
#pragma unroll
            for (int c = 0; c < RED_CHAINS; ++c) acc[c] = R::combine(acc[c], WordOps<R>::finish(wacc[c]));
        } else {
#pragma unroll
            for (int j = 0; j < RED_NV; ++j) {
                if (SHIFTED && j == RED_NV - 1 && last_slot_empty) continue;
                const int64_t e0 = gidx0 + head + ((int64_t)j * RED_THREADS + threadIdx.x) * EPV;
#pragma unroll
                for (int e = 0; e < EPV; ++e) acc[j % RED_CHAINS] = R::step(acc[j % RED_CHAINS], r[j].get(e), e0 + e);
            }
        }
        if (SHIFTED && threadIdx.x < EPV)
            acc[0] = R::step(acc[0], extra, gidx0 + leftover_pos<EPV>((int)threadIdx.x, head));
    } else {
        // ragged last block, strided or misaligned input: element loads, same fixed order
        for (int64_t i = threadIdx.x; i < cnt; i += RED_THREADS)
            acc[0] = R::step(acc[0], *reinterpret_cast<const In*>(base + i * stride), gidx0 + i);
    }

    Acc a = acc[0];
#pragma unroll
    for (int c = 1; c < RED_CHAINS; ++c) a = R::combine(a, acc[c]);
    a = block_fold<R, RED_THREADS>(a, smem);
    if (threadIdx.x == 0) st_partial(partials + k, a);
    if constexpr (FUSED) fused_tail<R, true>(partials, fa);
}

// ---- pass 1 of sum(a*b + c) ------------------------------------------------------
// The element stream fed to the SumR policy is u[i] = rn(rn(a[i]*b[i]) + c[i]) -- exactly the array
// np.add(np.multiply(a, b), c) would have stored -- cut into the same logical blocks and folded in
// the same order as reduce_blocks_kernel (vector j of a thread goes to accumulator j % 4, j
// ascending), so partial[k], and therefore the final sum, is bit-identical to the three-call chain
// while HBM is read once (12 B per float32 element instead of 28).  The three operands are
// loaded four vectors at a time (384 B in flight per thread).
template <class R, bool VEC, bool FUSED>
__global__ void __launch_bounds__(RED_THREADS, FUSED ? 3 : 1)   // one-launch form: room for the tail's call frame (at 64 registers the main loop spills)
muladd_sum_blocks_kernel(const typename R::In* a, const typename R::In* b, const typename R::In* c,
                         int64_t n, typename R::Acc* partials, const __grid_constant__ FusedArgs fa) {
    typedef typename R::In In;
    typedef typename R::Acc Acc;
    constexpr int64_t BE = block_elems<In>();
    constexpr int EPV = 32 / (int)sizeof(In);
    static_assert(RED_NV % RED_CHAINS == 0, "vector j must map to accumulator j % RED_CHAINS");
    __shared__ Acc smem[RED_THREADS / 32];

    const int64_t k = blockIdx.x;
    const int64_t start = k * BE;
    const int64_t cnt = (n - start) < BE ? (n - start) : BE;
    a += start; b += start; c += start;

    Acc acc[RED_CHAINS];
#pragma unroll
    for (int ch = 0; ch < RED_CHAINS; ++ch) acc[ch] = R::identity();

    if (VEC && cnt == BE) {
#pragma unroll 1
        for (int j0 = 0; j0 < RED_NV; j0 += RED_CHAINS) {
            Pack<In, EPV> ra[RED_CHAINS], rb[RED_CHAINS], rc[RED_CHAINS];
#pragma unroll
            for (int jj = 0; jj < RED_CHAINS; ++jj) {
                const int64_t e0 = ((int64_t)(j0 + jj) * RED_THREADS + threadIdx.x) * EPV;
                ra[jj] = ld_pack<In, EPV>(a + e0);
                rb[jj] = ld_pack<In, EPV>(b + e0);
                rc[jj] = ld_pack<In, EPV>(c + e0);
            }
#pragma unroll
            for (int jj = 0; jj < RED_CHAINS; ++jj)
#pragma unroll
                for (int e = 0; e < EPV; ++e)
                    acc[jj] = R::step(acc[jj], op_add<In>(op_mul<In>(ra[jj].get(e), rb[jj].get(e)), rc[jj].get(e)), 0);
        }
    } else {
        for (int64_t i = threadIdx.x; i < cnt; i += RED_THREADS)
            acc[0] = R::step(acc[0], op_add<In>(op_mul<In>(a[i], b[i]), c[i]), 0);
    }

    Acc t = acc[0];
#pragma unroll
    for (int ch = 1; ch < RED_CHAINS; ++ch) t = R::combine(t, acc[ch]);
    t = block_fold<R, RED_THREADS>(t, smem);
    if (threadIdx.x == 0) st_partial(partials + k, t);
    if constexpr (FUSED) fused_tail<R, true>(partials, fa);
}

// NaN-propagating extreme for phase A of the arg kernels
template <class T, bool IS_MAX> __device__ __forceinline__ T ext_nan(T a, T b) {
    if constexpr (std::is_same<T, float>::value) {
        float r;
        if (IS_MAX) asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
        else asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
        return r;
    } else if constexpr (is_fp<T>::value) {
        return (b != b) ? b : ((IS_MAX ? b > a : b < a) ? b : a);  // a NaN in `a` stays
    } else {
        return IS_MAX ? (b > a ? b : a) : (b < a ? b : a);
    }
}
template <class T, bool IS_MAX> struct ExtR {  // block fold of phase A
    typedef T In; typedef typename Widen<T>::type Acc;
    static __device__ __forceinline__ Acc combine(Acc a, Acc b) { return (Acc)ext_nan<Acc, IS_MAX>(a, b); }
};
struct IdxMinR {  // block fold of phase B
    typedef int In; typedef int Acc;
    static __device__ __forceinline__ int combine(int a, int b) { return a < b ? a : b; }
};

// arg-reduction pass 1: partial[k] = {extreme of block k, global index of its first occurrence}
template <class T, bool IS_MAX, int VEC, bool FUSED>
__global__ void __launch_bounds__(RED_THREADS, 3)
arg_blocks_kernel(const T* in, int64_t n, int64_t index_base, ArgPair<T>* partials, const __grid_constant__ FusedArgs fa) {
    typedef typename Widen<T>::type W;
    constexpr int64_t BE = block_elems<T>();
    constexpr int EPV = 32 / (int)sizeof(T);
    constexpr int NONE = 0x7fffffff;
    __shared__ W smem_v[RED_THREADS / 32];
    __shared__ int smem_i[RED_THREADS / 32];

    const int64_t k = blockIdx.x;
    const int64_t start = k * BE;
    const int64_t cnt = (n - start) < BE ? (n - start) : BE;
    const T* p = in + start;
    const bool full = VEC != RD_ELEMS && cnt == BE;

    Pack<T, EPV> r[RED_NV];
    W m = (W)p[0];
    constexpr bool SHIFTED = VEC == RD_SHIFTED;
    const int head = SHIFTED ? block_head<T, EPV>(p) : 0;   // see load_block
    const bool has_extra = SHIFTED && full && threadIdx.x < EPV;  // this thread also holds one leftover element
    T extra = T();
    // 1- and 2-byte integers: both phases a WORD at a time.  A 64 KiB block of int8 holds at most 256 different
    // values, so the block extreme sits in almost every thread's registers and phase B -- "usually one thread" for
    // the wide types -- runs everywhere: at three instructions per byte the kernel was compute-bound (31 % of 8 TB/s).
    constexpr bool WORDWISE = !is_fp<T>::value && sizeof(T) <= 2;
    if (full) {
        load_block<T, EPV, SHIFTED>(p, head, r, extra);
        // (thread 255's duplicated last pack is harmless here: phase A is idempotent, phase B skips it)
        if constexpr (WORDWISE) {
            typedef MinMaxWord<T, IS_MAX> MW;
            typename MW::W acc = MW::seed(p[0]);
#pragma unroll
            for (int j = 0; j < RED_NV; ++j)
#pragma unroll
                for (int w = 0; w < Pack<T, EPV>::WORDS; ++w) acc = MW::step(acc, r[j].w[w]);
            m = ext_nan<W, IS_MAX>(m, (W)MW::finish(acc));
        } else {
#pragma unroll
            for (int j = 0; j < RED_NV; ++j)
#pragma unroll
                for (int e = 0; e < EPV; ++e) m = ext_nan<W, IS_MAX>(m, (W)r[j].get(e));
        }
        if (has_extra) m = ext_nan<W, IS_MAX>(m, (W)extra);
    } else {
        for (int64_t i = threadIdx.x; i < cnt; i += RED_THREADS) m = ext_nan<W, IS_MAX>(m, (W)p[i]);
    }
    // block extreme with ONE barrier: shuffle tree per warp, leaders to shared memory, then every
    // thread folds the eight warp values itself (same order everywhere -> same value everywhere)
    const W m_thread = m;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = ext_nan<W, IS_MAX>(m, Shfl<W>::down(m, o));
    if ((threadIdx.x & 31) == 0) smem_v[threadIdx.x >> 5] = m;
    __syncthreads();
    W mb = smem_v[0];
#pragma unroll
    for (int w = 1; w < RED_THREADS / 32; ++w) mb = ext_nan<W, IS_MAX>(mb, smem_v[w]);
    bool want_nan = false;
    if constexpr (is_fp<T>::value) want_nan = mb != mb;

    // phase B only in the threads whose own extreme IS the block extreme (usually one thread of one
    // warp): they scan their registers for its first occurrence
    int local = NONE;
    bool mine = want_nan ? false : (m_thread == mb);
    if constexpr (is_fp<T>::value) mine = want_nan ? (m_thread != m_thread) : mine;
    if (mine) {
        if (full) {
            int c = NONE;  // scan downwards so the smallest matching position survives
            if (want_nan) {
                if constexpr (is_fp<T>::value) {
#pragma unroll
                    for (int j = RED_NV - 1; j >= 0; --j)
#pragma unroll
                        for (int e = EPV - 1; e >= 0; --e) {
                            const W x = (W)r[j].get(e);
                            const bool real = j < RED_NV - 1 || !(SHIFTED && threadIdx.x == RED_THREADS - 1);
                            c = (x != x && real) ? (j * RED_THREADS * EPV + e) : c;
                        }
                }
            } else if constexpr (WORDWISE) {
                // zero-lane detector on word ^ splat(extreme): the LOWEST flagged lane of a word is exact (a borrow can
                // only flag lanes above a true match), and the lowest lane is the first element (little endian)
                constexpr int WORDS = Pack<T, EPV>::WORDS, EPW = 4 / (int)sizeof(T);
                constexpr uint32_t ONES = sizeof(T) == 1 ? 0x01010101u : 0x00010001u;
                constexpr uint32_t HIGH = sizeof(T) == 1 ? 0x80808080u : 0x80008000u;
                const uint32_t splat = splat_word<T>((T)mb);
                uint32_t cz = 0u;
                int cw = 0;
#pragma unroll
                for (int j = RED_NV - 1; j >= 0; --j)
#pragma unroll
                    for (int w = WORDS - 1; w >= 0; --w) {
                        const uint32_t t = r[j].w[w] ^ splat;
                        uint32_t z = (t - ONES) & ~t & HIGH;
                        if (j == RED_NV - 1 && SHIFTED && threadIdx.x == RED_THREADS - 1) z = 0u;
                        if (z) { cz = z; cw = j * WORDS + w; }
                    }
                if (cz) c = (cw / WORDS) * (RED_THREADS * EPV) + (cw % WORDS) * EPW + ((__ffs((int)cz) - 1) >> (sizeof(T) == 1 ? 3 : 4));
            } else {
#pragma unroll
                for (int j = RED_NV - 1; j >= 0; --j)
#pragma unroll
                    for (int e = EPV - 1; e >= 0; --e) {
                        const W x = (W)r[j].get(e);
                        const bool real = j < RED_NV - 1 || !(SHIFTED && threadIdx.x == RED_THREADS - 1);
                        c = (x == mb && real) ? (j * RED_THREADS * EPV + e) : c;
                    }
            }
            if (c != NONE) local = c + (int)threadIdx.x * EPV + head;
            if (has_extra) {   // the leftover elements sit at the two ends of the block
                const W x = (W)extra;
                bool hit = x == mb;
                if constexpr (is_fp<T>::value) hit = want_nan ? (x != x) : hit;
                const int pos = leftover_pos<EPV>((int)threadIdx.x, head);
                if (hit && pos < local) local = pos;
            }
        } else {
            for (int64_t i = threadIdx.x; i < cnt; i += RED_THREADS) {
                const W x = (W)p[i];
                const bool hit = want_nan ? (x != x) : (x == mb);
                if (hit && (int)i < local) local = (int)i;
            }
        }
    }
    local = block_fold<IdxMinR, RED_THREADS>(local, smem_i);
    if (threadIdx.x == 0) {
        ArgPair<T> out;
        out.v = (T)mb;
        out.idx = start + index_base + local;
        st_partial(partials + k, out);
    }
    if constexpr (FUSED) fused_tail<ArgR<T, IS_MAX>, false>(partials, fa);
}

// ---- pass 2 --------------------------------------------------------------------------------------
// The two-launch building block (pncu_reduce_finish / pncu_argminmax_finish: the multi-GPU dispatcher's
// "gather, then fold" and the shard-invariance tests).  VALUE: write R::result into `out` (reductions);
// else write the index (arg-reductions).
template <class R, bool VALUE>
__global__ void __launch_bounds__(FIN_THREADS)
reduce_finish_kernel(const typename R::Acc* partials, int64_t count, typename R::Acc* gscratch, void* out, const void* startVal) {
    typedef typename R::Acc Acc;
    __shared__ Acc smem[FIN_THREADS / 32];
    Acc a = fold_two_level<R, FIN_THREADS>(partials, count, gscratch, smem);
    if (threadIdx.x == 0) store_result<R, VALUE>(a, out, startVal);
}

// ---- pass 2 with the cross-GPU exchange fused in ---------------------------------------------------------
// G participants (GPUs), each with `local_count` block partials from its own pass 1.  One kernel per
// participant does the whole exchange and the fold:
//   CTA (s, d), s < EXCH_SLICES, d < G: copies slice s of the local partials into participant d's slot table
//       at slot_base (plain stores into peer memory over NVLink; d == self is the local table), then one
//       system-scope fence and one arrival count on d's counter;
//   CTA (0, self) then waits ON THE DEVICE until all G * EXCH_SLICES slices of this call have arrived in its
//       own table and folds the complete table exactly like reduce_finish_kernel (same order, same bits).
// No collective call, no host synchronisation, no extra launch compared with a single GPU; the payload is
// 4-16 bytes per 64 KiB of input.  All G * EXCH_SLICES CTAs (<= 64) are co-resident, so the waiting CTA
// cannot starve the pushing ones.  The wait is bounded (EXCH_WAIT_NS): when a participant never arrives the
// table is NOT folded and `out` is NOT written; the kernel counts the timeout (pncu_exchange_timeouts) and
// writes PNCU_FUSED_TIMEOUT_BIT into `status` so that the host raises instead of using a partial result.
// (The one-launch form of the same exchange is fused_finish above; this kernel remains for callers that
// produce their partials separately.)
constexpr int EXCH_SLICES = 8;
struct PeerOut {
    int n, self;
    long long slot_base;
    void* slots[PNCU_MAX_PEERS];
    unsigned long long* arrived[PNCU_MAX_PEERS];
};

template <class R, bool VALUE>
__global__ void __launch_bounds__(FIN_THREADS)
exchange_finish_kernel(const typename R::Acc* local, int64_t local_count, const PeerOut po, int64_t total_count,
                       typename R::Acc* gscratch, void* out, const void* startVal, unsigned long long expected,
                       unsigned long long* status) {
    typedef typename R::Acc Acc;
    __shared__ Acc smem[FIN_THREADS / 32];
    __shared__ int s_ok;
    const int s = blockIdx.x, d = blockIdx.y;
    // ---- push slice s to participant d
    {
        const int64_t per = (local_count + EXCH_SLICES - 1) / EXCH_SLICES;
        const int64_t lo = (int64_t)s * per;
        const int64_t hi = lo + per < local_count ? lo + per : local_count;
        Acc* dst = reinterpret_cast<Acc*>(po.slots[d]) + po.slot_base;
        for (int64_t i = lo + threadIdx.x; i < hi; i += FIN_THREADS) dst[i] = local[i];
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence_system();
            atomicAdd_system(po.arrived[d], 1ULL);
        }
    }
    if (s != 0 || d != po.self) return;
    // ---- wait for everybody's slices, then fold the whole table
    if (threadIdx.x == 0) s_ok = wait_arrivals(po.arrived[po.self], expected) ? 1 : 0;
    __syncthreads();
    if (!s_ok) {
        if (threadIdx.x == 0 && status) atomicOr(status, PNCU_FUSED_TIMEOUT_BIT);
        return;
    }
    Acc a = fold_two_level<R, FIN_THREADS>(reinterpret_cast<const Acc*>(po.slots[po.self]), total_count, gscratch, smem);
    if (threadIdx.x == 0) store_result<R, VALUE>(a, out, startVal);
}

// ---- host launchers ---------------------------------------------------------------------------------
template <class In> inline int64_t nblocks_for(int64_t len) { return (len + block_elems<In>() - 1) / block_elems<In>(); }

static const FusedArgs kNoFuse = {};
// PNUMPY_CUDA_GROUP_TICKETS=0 keeps the classic tail (the last CTA runs both levels of pass 2)
inline bool group_tickets_enabled() {
    static const bool on = [] { const char* e = getenv("PNUMPY_CUDA_GROUP_TICKETS"); return !(e && *e == '0'); }();
    return on;
}
// does this participant of a one-launch reduction fold (and therefore need `out`)?
inline bool consumes(const pncu_fused_exchange* x) { return !x || x->n <= 1 || x->all || x->self == x->root; }

// FusedArgs and the place pass 1 writes its partials for a ONE-LAUNCH reduction of `nb` blocks.
// x == NULL (or x->n == 1): single GPU, partials in the stream's scratch.  x->n > 1: see pncu_fused_exchange.
int prepare_fused(const pncu_fused_exchange* x, size_t acc_bytes, int64_t nb, void* out, const void* startVal,
                  cudaStream_t st, FusedArgs* fa, void** partials) {
    void* ws = NULL;
    unsigned* ticket = NULL;
    unsigned* gtickets = NULL;
    // scratch layout: [nb slots of this launch][group partials of the largest table this GPU may fold itself]
    const int64_t fold_slots = x && x->n > 1 && x->total_count > nb ? x->total_count : nb;
    // group tickets: whenever the local groups ARE groups of the whole table (single GPU, or shards that start at
    // whole groups) and there are enough of them to matter
    const bool group_mode = group_tickets_enabled() && nb >= 4 * FOLD_GROUP && (!x || x->n <= 1 || x->group_offset > 0);
    int rc = get_workspace(st, (size_t)(nb + groups_for(fold_slots)) * acc_bytes, &ws, &ticket,
                           group_mode ? (size_t)groups_for(nb) : 0, group_mode ? &gtickets : nullptr);
    if (rc) return rc;
    *fa = FusedArgs();
    fa->ticket = ticket;
    fa->gtickets = gtickets;
    fa->n = 1;
    fa->out = out;
    fa->startVal = startVal;
    fa->gscratch = (char*)ws + (size_t)nb * acc_bytes;
    *partials = ws;
    if (!x) return 0;
    fa->done_flag = x->done_flag;
    fa->done_value = x->done_value;
    if (x->n == 1) return 0;
    if (x->n < 1 || x->n > PNCU_MAX_PEERS || x->self < 0 || x->self >= x->n || x->root < 0 || x->root >= x->n) {
        set_error("fused exchange: n must be 1..%d, 0 <= self, root < n", PNCU_MAX_PEERS);
        return PNCU_ERR_ARGUMENT;
    }
    if (x->slot_base < 0 || x->slot_base + nb > x->total_count) {
        set_error("fused exchange: slots [%lld, %lld) do not fit a table of %lld", (long long)x->slot_base,
                  (long long)(x->slot_base + nb), (long long)x->total_count);
        return PNCU_ERR_ARGUMENT;
    }
    fa->n = x->n; fa->self = x->self; fa->all = x->all ? 1 : 0; fa->root = x->root;
    fa->slot_base = x->slot_base; fa->total_count = x->total_count; fa->expected = x->expected;
    for (int d = 0; d < x->n; ++d) {
        const bool used = x->all || d == x->root;
        if (used && (!x->slots[d] || !x->arrived[d])) { set_error("fused exchange: null table for participant %d", d); return PNCU_ERR_ARGUMENT; }
        fa->slots[d] = x->slots[d];
        fa->arrived[d] = x->arrived[d];
    }
    if (x->group_offset > 0) {
        if (x->slot_base % FOLD_GROUP || x->group_offset % 16 || (size_t)x->group_offset < (size_t)x->total_count * acc_bytes) {
            set_error("fused exchange: group_offset %lld needs slot_base %lld to be a multiple of %d and the group table to lie behind "
                      "the %lld slots, 16-byte aligned", (long long)x->group_offset, (long long)x->slot_base, FOLD_GROUP, (long long)x->total_count);
            return PNCU_ERR_ARGUMENT;
        }
        fa->group_off = x->group_offset;
    }
    if (x->all || x->self == x->root) *partials = (char*)x->slots[x->self] + (size_t)x->slot_base * acc_bytes;
    return 0;
}

template <class R>
int launch_partials(const void* in, int64_t len, int64_t stride, int64_t index_base, void* partials,
                    cudaStream_t st, const FusedArgs* fa = nullptr) {
    typedef typename R::In In;
    typedef typename R::Acc Acc;
    const int64_t nb = nblocks_for<In>(len);
    if (nb > 0x7fffffff) { set_error("length %lld exceeds the grid limit", (long long)len); return PNCU_ERR_ARGUMENT; }
    const int mode = read_mode(in, stride == (int64_t)sizeof(In));
    const char* p = (const char*)in;
    const unsigned g = (unsigned)nb;
#define PNCU_LAUNCH_RED(MODE, FUSEDFLAG, ARGS) \
    reduce_blocks_kernel<R, MODE, FUSEDFLAG><<<g, RED_THREADS, 0, st>>>(p, len, stride, index_base, (Acc*)partials, ARGS)
    if (fa) {
        if (mode == RD_VEC) PNCU_LAUNCH_RED(RD_VEC, true, *fa);
        else if (mode == RD_SHIFTED) PNCU_LAUNCH_RED(RD_SHIFTED, true, *fa);
        else PNCU_LAUNCH_RED(RD_ELEMS, true, *fa);
    } else {
        if (mode == RD_VEC) PNCU_LAUNCH_RED(RD_VEC, false, kNoFuse);
        else if (mode == RD_SHIFTED) PNCU_LAUNCH_RED(RD_SHIFTED, false, kNoFuse);
        else PNCU_LAUNCH_RED(RD_ELEMS, false, kNoFuse);
    }
#undef PNCU_LAUNCH_RED
    return after_launch("reduce_blocks_kernel");
}

template <class T, bool IS_MAX>
int launch_arg_partials(const void* in, int64_t len, int64_t /*stride*/, int64_t index_base, void* partials,
                        cudaStream_t st, const FusedArgs* fa = nullptr) {
    const int64_t nb = nblocks_for<T>(len);
    if (nb > 0x7fffffff) { set_error("length %lld exceeds the grid limit", (long long)len); return PNCU_ERR_ARGUMENT; }
    const int mode = read_mode(in, true);   // contiguous by contract
    const unsigned g = (unsigned)nb;
    ArgPair<T>* pp = (ArgPair<T>*)partials;
#define PNCU_LAUNCH_ARG(MODE, FUSEDFLAG, ARGS) \
    arg_blocks_kernel<T, IS_MAX, MODE, FUSEDFLAG><<<g, RED_THREADS, 0, st>>>((const T*)in, len, index_base, pp, ARGS)
    if (fa) {
        if (mode == RD_VEC) PNCU_LAUNCH_ARG(RD_VEC, true, *fa);
        else PNCU_LAUNCH_ARG(RD_SHIFTED, true, *fa);
    } else {
        if (mode == RD_VEC) PNCU_LAUNCH_ARG(RD_VEC, false, kNoFuse);
        else PNCU_LAUNCH_ARG(RD_SHIFTED, false, kNoFuse);
    }
#undef PNCU_LAUNCH_ARG
    return after_launch("arg_blocks_kernel");
}

template <class R>
int check_reduce_args(const void* in, int64_t stride) {
    typedef typename R::In In;
    if (!in) { set_error("null input"); return PNCU_ERR_ARGUMENT; }
    if (stride == 0) { set_error("reduction over a stride-0 operand has no GPU body"); return PNCU_ERR_UNSUPPORTED; }
    if (((uintptr_t)in | (uint64_t)stride) % sizeof(In)) {
        set_error("operand pointer or stride is not a multiple of the item size");
        return PNCU_ERR_UNSUPPORTED;
    }
    return 0;
}

// the complete reduction in ONE launch (optionally with the cross-GPU exchange)
template <class R>
int launch_reduce_fused(const void* in, int64_t len, int64_t stride, int64_t /*index_base*/, const pncu_fused_exchange* x,
                        void* out, const void* startVal, cudaStream_t st) {
    typedef typename R::In In;
    int rc;
    if (!out && consumes(x)) { set_error("null output"); return PNCU_ERR_ARGUMENT; }
    if (len <= 0) { set_error("one-launch reduction of an empty shard"); return PNCU_ERR_ARGUMENT; }
    if ((rc = check_reduce_args<R>(in, stride))) return rc;
    FusedArgs fa;
    void* partials = NULL;
    if ((rc = prepare_fused(x, sizeof(typename R::Acc), nblocks_for<In>(len), out, startVal, st, &fa, &partials))) return rc;
    return launch_partials<R>(in, len, stride, 0, partials, st, &fa);
}

// REDUCE_FUNC
template <class R>
int launch_reduce(const void* in, void* out, const void* startVal, int64_t len, int64_t stride,
                  pncu_stream_t stream) {
    typedef typename R::In In;
    int rc = ensure_init();
    if (rc) return rc;
    cudaStream_t st = as_stream(stream);
    if (!out) { set_error("null output"); return PNCU_ERR_ARGUMENT; }
    if (len <= 0) {
        if (startVal && startVal != out) PNCU_CUDA(cudaMemcpyAsync(out, startVal, sizeof(In), cudaMemcpyDeviceToDevice, st));
        return 0;
    }
    return launch_reduce_fused<R>(in, len, stride, 0, NULL, out, startVal, st);
}

// sum(a*b + c): pass 1 (fused multiply-add) with the block structure of the ADD reduction
template <class R>
int launch_muladd_partials(const void* a, const void* b, const void* c, int64_t len, void* partials, cudaStream_t st,
                           const FusedArgs* fa = nullptr) {
    typedef typename R::In In;
    typedef typename R::Acc Acc;
    if (!a || !b || !c || !partials) { set_error("null operand"); return PNCU_ERR_ARGUMENT; }
    if (((uintptr_t)a | (uintptr_t)b | (uintptr_t)c) % sizeof(In)) {
        set_error("operand pointer is not a multiple of the item size");
        return PNCU_ERR_UNSUPPORTED;
    }
    const int64_t nb = nblocks_for<In>(len);
    if (nb > 0x7fffffff) { set_error("length %lld exceeds the grid limit", (long long)len); return PNCU_ERR_ARGUMENT; }
    const bool vec = (((uintptr_t)a | (uintptr_t)b | (uintptr_t)c) % 32) == 0;
    const unsigned g = (unsigned)nb;
    const In *pa = (const In*)a, *pb = (const In*)b, *pc = (const In*)c;
    if (fa) {
        if (vec) muladd_sum_blocks_kernel<R, true, true><<<g, RED_THREADS, 0, st>>>(pa, pb, pc, len, (Acc*)partials, *fa);
        else muladd_sum_blocks_kernel<R, false, true><<<g, RED_THREADS, 0, st>>>(pa, pb, pc, len, (Acc*)partials, *fa);
    } else {
        if (vec) muladd_sum_blocks_kernel<R, true, false><<<g, RED_THREADS, 0, st>>>(pa, pb, pc, len, (Acc*)partials, kNoFuse);
        else muladd_sum_blocks_kernel<R, false, false><<<g, RED_THREADS, 0, st>>>(pa, pb, pc, len, (Acc*)partials, kNoFuse);
    }
    return after_launch("muladd_sum_blocks_kernel");
}

template <class R>
int launch_muladd_sum_fused(const void* a, const void* b, const void* c, int64_t len, const pncu_fused_exchange* x, void* out,
                            cudaStream_t st) {
    typedef typename R::In In;
    int rc;
    if (!out && consumes(x)) { set_error("null output"); return PNCU_ERR_ARGUMENT; }
    if (len <= 0) { set_error("one-launch reduction of an empty shard"); return PNCU_ERR_ARGUMENT; }
    FusedArgs fa;
    void* partials = NULL;
    if ((rc = prepare_fused(x, sizeof(typename R::Acc), nblocks_for<In>(len), out, NULL, st, &fa, &partials))) return rc;
    return launch_muladd_partials<R>(a, b, c, len, partials, st, &fa);
}

template <class R>
int launch_muladd_sum(const void* a, const void* b, const void* c, void* out, int64_t len, pncu_stream_t stream) {
    typedef typename R::In In;
    int rc = ensure_init();
    if (rc) return rc;
    cudaStream_t st = as_stream(stream);
    if (!out) { set_error("null output"); return PNCU_ERR_ARGUMENT; }
    if (len <= 0) { PNCU_CUDA(cudaMemsetAsync(out, 0, sizeof(In), st)); return 0; }
    return launch_muladd_sum_fused<R>(a, b, c, len, NULL, out, st);
}

struct MulAddSumOps {
    pncu_three_reduce_fn full;
    int (*partials)(const void*, const void*, const void*, int64_t, void*, cudaStream_t, const FusedArgs*);
    int (*fused)(const void*, const void*, const void*, int64_t, const pncu_fused_exchange*, void*, cudaStream_t);
};
template <class R> MulAddSumOps make_muladd_ops() {
    MulAddSumOps o;
    o.full = launch_muladd_sum<R>;
    o.partials = launch_muladd_partials<R>;
    o.fused = launch_muladd_sum_fused<R>;
    return o;
}
bool lookup_muladd_sum(int t, MulAddSumOps* o) {
    switch (t) {  // the accumulator types of lookup_reduce(PNCU_ADD, t)
        case PNCU_INT32: *o = make_muladd_ops<SumR<int32_t, unsigned>>(); return true;
        case PNCU_UINT32: *o = make_muladd_ops<SumR<uint32_t, unsigned>>(); return true;
        case PNCU_INT64: *o = make_muladd_ops<SumR<int64_t, unsigned long long>>(); return true;
        case PNCU_UINT64: *o = make_muladd_ops<SumR<uint64_t, unsigned long long>>(); return true;
        case PNCU_FLOAT: *o = make_muladd_ops<SumR<float, double>>(); return true;
        case PNCU_DOUBLE: *o = make_muladd_ops<SumR<double, double>>(); return true;
    }
    return false;
}

template <class R>
int launch_arg_fused(const void* in, int64_t len, int64_t /*stride*/, int64_t index_base, const pncu_fused_exchange* x,
                     void* out_index, const void* /*startVal*/, cudaStream_t st) {
    int rc;
    if (!out_index && consumes(x)) { set_error("null output"); return PNCU_ERR_ARGUMENT; }
    if (len <= 0) { set_error("arg-reduction of an empty sequence"); return PNCU_ERR_ARGUMENT; }
    if ((rc = check_reduce_args<R>(in, sizeof(typename R::In)))) return rc;
    FusedArgs fa;
    void* partials = NULL;
    if ((rc = prepare_fused(x, sizeof(typename R::Acc), nblocks_for<typename R::In>(len), out_index, NULL, st, &fa, &partials))) return rc;
    return launch_arg_partials<typename R::In, R::kIsMax>(in, len, sizeof(typename R::In), index_base, partials, st, &fa);
}

template <class R>
int launch_argminmax(const void* in, int64_t len, int64_t* out_index, pncu_stream_t stream) {
    int rc = ensure_init();
    if (rc) return rc;
    return launch_arg_fused<R>(in, len, sizeof(typename R::In), 0, NULL, out_index, NULL, as_stream(stream));
}

// type-erased entry for the building blocks
struct ReduceOps {
    int partial_bytes;
    pncu_reduce_fn full;
    int (*partials)(const void*, int64_t, int64_t, int64_t, void*, cudaStream_t, const FusedArgs*);
    int (*finish)(const void*, int64_t, void*, const void*, cudaStream_t);
    int (*exchange_finish)(const void*, int64_t, const PeerOut*, int64_t, void*, const void*, unsigned long long,
                           unsigned long long*, cudaStream_t);
    int (*fused)(const void*, int64_t, int64_t, int64_t, const pncu_fused_exchange*, void*, const void*, cudaStream_t);
};
template <class R, bool VALUE>
int finish_thunk(const void* partials, int64_t count, void* out, const void* startVal, cudaStream_t st) {
    void* gs = NULL;   // level-A output: this stream's scratch (the caller's partials live elsewhere)
    int rc = get_workspace(st, (size_t)groups_for(count) * sizeof(typename R::Acc), &gs);
    if (rc) return rc;
    reduce_finish_kernel<R, VALUE><<<1, FIN_THREADS, 0, st>>>((const typename R::Acc*)partials, count, (typename R::Acc*)gs, out, startVal);
    return after_launch("reduce_finish_kernel");
}
template <class R, bool VALUE>
int exchange_finish_thunk(const void* local, int64_t local_count, const PeerOut* po, int64_t total_count, void* out,
                          const void* startVal, unsigned long long expected, unsigned long long* status, cudaStream_t st) {
    void* gs = NULL;
    int rc = get_workspace(st, (size_t)groups_for(total_count) * sizeof(typename R::Acc), &gs);
    if (rc) return rc;
    exchange_finish_kernel<R, VALUE><<<dim3(EXCH_SLICES, po->n), FIN_THREADS, 0, st>>>(
        (const typename R::Acc*)local, local_count, *po, total_count, (typename R::Acc*)gs, out, startVal, expected, status);
    return after_launch("exchange_finish_kernel");
}
template <class R>
ReduceOps make_ops() {
    ReduceOps o;
    o.partial_bytes = (int)sizeof(typename R::Acc);
    o.full = launch_reduce<R>;
    o.partials = launch_partials<R>;
    o.finish = finish_thunk<R, true>;
    o.exchange_finish = exchange_finish_thunk<R, true>;
    o.fused = launch_reduce_fused<R>;
    return o;
}
template <class R>
ReduceOps make_arg_ops() {
    ReduceOps o;
    o.partial_bytes = (int)sizeof(typename R::Acc);
    o.full = NULL;
    o.partials = launch_arg_partials<typename R::In, R::kIsMax>;
    o.finish = finish_thunk<R, false>;
    o.exchange_finish = exchange_finish_thunk<R, false>;
    o.fused = launch_arg_fused<R>;
    return o;
}

bool lookup_reduce(int func, int t, ReduceOps* o) {
    switch (func) {
        case PNCU_ADD:
            switch (t) {
                case PNCU_BOOL: *o = make_ops<AnyR>(); return true;
                case PNCU_INT8: *o = make_ops<SumR<int8_t, unsigned>>(); return true;
                case PNCU_UINT8: *o = make_ops<SumR<uint8_t, unsigned>>(); return true;
                case PNCU_INT16: *o = make_ops<SumR<int16_t, unsigned>>(); return true;
                case PNCU_UINT16: *o = make_ops<SumR<uint16_t, unsigned>>(); return true;
                case PNCU_INT32: *o = make_ops<SumR<int32_t, unsigned>>(); return true;
                case PNCU_UINT32: *o = make_ops<SumR<uint32_t, unsigned>>(); return true;
                case PNCU_INT64: *o = make_ops<SumR<int64_t, unsigned long long>>(); return true;
                case PNCU_UINT64: *o = make_ops<SumR<uint64_t, unsigned long long>>(); return true;
                case PNCU_FLOAT: *o = make_ops<SumR<float, double>>(); return true;
                case PNCU_DOUBLE: *o = make_ops<SumR<double, double>>(); return true;
            }
            return false;
        case PNCU_MUL:
            switch (t) {
                case PNCU_BOOL: *o = make_ops<AllR>(); return true;
                case PNCU_INT8: *o = make_ops<ProdR<int8_t, unsigned>>(); return true;
                case PNCU_UINT8: *o = make_ops<ProdR<uint8_t, unsigned>>(); return true;
                case PNCU_INT16: *o = make_ops<ProdR<int16_t, unsigned>>(); return true;
                case PNCU_UINT16: *o = make_ops<ProdR<uint16_t, unsigned>>(); return true;
                case PNCU_INT32: *o = make_ops<ProdR<int32_t, unsigned>>(); return true;
                case PNCU_UINT32: *o = make_ops<ProdR<uint32_t, unsigned>>(); return true;
                case PNCU_INT64: *o = make_ops<ProdR<int64_t, unsigned long long>>(); return true;
                case PNCU_UINT64: *o = make_ops<ProdR<uint64_t, unsigned long long>>(); return true;
                case PNCU_FLOAT: *o = make_ops<ProdR<float, double>>(); return true;
                case PNCU_DOUBLE: *o = make_ops<ProdR<double, double>>(); return true;
            }
            return false;
#define PNCU_MINMAX_CASES(RR)                                                  \
    switch (t) {                                                               \
        case PNCU_BOOL: case PNCU_UINT8: *o = make_ops<RR<uint8_t>>(); return true; \
        case PNCU_INT8: *o = make_ops<RR<int8_t>>(); return true;              \
        case PNCU_INT16: *o = make_ops<RR<int16_t>>(); return true;            \
        case PNCU_UINT16: *o = make_ops<RR<uint16_t>>(); return true;          \
        case PNCU_INT32: *o = make_ops<RR<int32_t>>(); return true;            \
        case PNCU_UINT32: *o = make_ops<RR<uint32_t>>(); return true;          \
        case PNCU_INT64: *o = make_ops<RR<int64_t>>(); return true;            \
        case PNCU_UINT64: *o = make_ops<RR<uint64_t>>(); return true;          \
        case PNCU_FLOAT: *o = make_ops<RR<float>>(); return true;              \
        case PNCU_DOUBLE: *o = make_ops<RR<double>>(); return true;            \
    }                                                                          \
    return false;
        case PNCU_MIN: PNCU_MINMAX_CASES(MinR)
        case PNCU_MAX: PNCU_MINMAX_CASES(MaxR)
#undef PNCU_MINMAX_CASES
        // bool any()/all() arrive as logical_or / logical_and reductions (the reference has no fast
        // body for them and threads NumPy's loop; the OR / AND folds above are the same thing)
        case PNCU_LOGICAL_OR: case PNCU_BITWISE_OR:
            if (t == PNCU_BOOL) { *o = make_ops<AnyR>(); return true; }
            return false;
        case PNCU_LOGICAL_AND: case PNCU_BITWISE_AND:
            if (t == PNCU_BOOL) { *o = make_ops<AllR>(); return true; }
            return false;
    }
    return false;
}

template <bool IS_MAX>
bool lookup_arg_t(int t, ReduceOps* o) {
    switch (t) {
        case PNCU_BOOL: case PNCU_UINT8: *o = make_arg_ops<ArgR<uint8_t, IS_MAX>>(); return true;
        case PNCU_INT8: *o = make_arg_ops<ArgR<int8_t, IS_MAX>>(); return true;
        case PNCU_INT16: *o = make_arg_ops<ArgR<int16_t, IS_MAX>>(); return true;
        case PNCU_UINT16: *o = make_arg_ops<ArgR<uint16_t, IS_MAX>>(); return true;
        case PNCU_INT32: *o = make_arg_ops<ArgR<int32_t, IS_MAX>>(); return true;
        case PNCU_UINT32: *o = make_arg_ops<ArgR<uint32_t, IS_MAX>>(); return true;
        case PNCU_INT64: *o = make_arg_ops<ArgR<int64_t, IS_MAX>>(); return true;
        case PNCU_UINT64: *o = make_arg_ops<ArgR<uint64_t, IS_MAX>>(); return true;
        case PNCU_FLOAT: *o = make_arg_ops<ArgR<float, IS_MAX>>(); return true;
        case PNCU_DOUBLE: *o = make_arg_ops<ArgR<double, IS_MAX>>(); return true;
    }
    return false;
}
bool lookup_arg(int kind, int t, ReduceOps* o) {
    if (kind == PNCU_ARGMAX) return lookup_arg_t<true>(t, o);
    if (kind == PNCU_ARGMIN) return lookup_arg_t<false>(t, o);
    return false;
}

template <bool IS_MAX>
pncu_argminmax_fn pick_arg(int t) {
    switch (t) {
        case PNCU_BOOL: case PNCU_UINT8: return launch_argminmax<ArgR<uint8_t, IS_MAX>>;
        case PNCU_INT8: return launch_argminmax<ArgR<int8_t, IS_MAX>>;
        case PNCU_INT16: return launch_argminmax<ArgR<int16_t, IS_MAX>>;
        case PNCU_UINT16: return launch_argminmax<ArgR<uint16_t, IS_MAX>>;
        case PNCU_INT32: return launch_argminmax<ArgR<int32_t, IS_MAX>>;
        case PNCU_UINT32: return launch_argminmax<ArgR<uint32_t, IS_MAX>>;
        case PNCU_INT64: return launch_argminmax<ArgR<int64_t, IS_MAX>>;
        case PNCU_UINT64: return launch_argminmax<ArgR<uint64_t, IS_MAX>>;
        case PNCU_FLOAT: return launch_argminmax<ArgR<float, IS_MAX>>;
        case PNCU_DOUBLE: return launch_argminmax<ArgR<double, IS_MAX>>;
    }
    return NULL;
}

}  // namespace

extern "C" pncu_reduce_fn pncu_GetReduceMathOpFast(int func, int t) {
    ReduceOps o;
    return lookup_reduce(func, t, &o) ? o.full : NULL;
}

extern "C" pncu_argminmax_fn pncu_GetArgMinMaxOpFast(int kind, int t) {
    if (kind == PNCU_ARGMAX) return pick_arg<true>(t);
    if (kind == PNCU_ARGMIN) return pick_arg<false>(t);
    return NULL;
}

extern "C" int64_t pncu_reduce_block_elems(void) { return RED_ALIGN_ELEMS; }
extern "C" int pncu_reduce_group_slots(void) { return FOLD_GROUP; }

extern "C" int64_t pncu_reduce_partial_count(int t, int64_t len) {
    const int isz = pncu_itemsize(t);
    if (!isz || len <= 0) return 0;
    const int64_t be = RED_BLOCK_BYTES / isz;
    return (len + be - 1) / be;
}

extern "C" int pncu_reduce_partial_bytes(int func, int t) {
    ReduceOps o;
    return lookup_reduce(func, t, &o) ? o.partial_bytes : 0;
}

extern "C" int pncu_reduce_partials(int func, int t, const void* in, int64_t len, void* partials,
                                    pncu_stream_t stream) {
    ReduceOps o;
    if (!lookup_reduce(func, t, &o)) { set_error("no reduction kernel for func %d type %d", func, t); return PNCU_ERR_UNSUPPORTED; }
    int rc = ensure_init();
    if (rc) return rc;
    if (len <= 0) return 0;
    if (!in || !partials) { set_error("null operand"); return PNCU_ERR_ARGUMENT; }
    return o.partials(in, len, pncu_itemsize(t), 0, partials, as_stream(stream), nullptr);
}

extern "C" int pncu_reduce_finish(int func, int t, const void* partials, int64_t count, void* out,
                                  const void* startVal, pncu_stream_t stream) {
    ReduceOps o;
    if (!lookup_reduce(func, t, &o)) { set_error("no reduction kernel for func %d type %d", func, t); return PNCU_ERR_UNSUPPORTED; }
    int rc = ensure_init();
    if (rc) return rc;
    if (!out) { set_error("null output"); return PNCU_ERR_ARGUMENT; }
    if (count <= 0) {
        if (startVal && startVal != out)
            PNCU_CUDA(cudaMemcpyAsync(out, startVal, pncu_itemsize(t), cudaMemcpyDeviceToDevice, as_stream(stream)));
        return 0;
    }
    return o.finish(partials, count, out, startVal, as_stream(stream));
}

extern "C" int pncu_argminmax_partials(int kind, int t, const void* in, int64_t len, int64_t index_base,
                                       void* partials, pncu_stream_t stream) {
    ReduceOps o;
    if (!lookup_arg(kind, t, &o)) { set_error("no arg-reduction kernel for kind %d type %d", kind, t); return PNCU_ERR_UNSUPPORTED; }
    int rc = ensure_init();
    if (rc) return rc;
    if (len <= 0) return 0;
    if (!in || !partials) { set_error("null operand"); return PNCU_ERR_ARGUMENT; }
    return o.partials(in, len, pncu_itemsize(t), index_base, partials, as_stream(stream), nullptr);
}

extern "C" int pncu_argminmax_finish(int kind, int t, const void* partials, int64_t count,
                                     int64_t* out_index, pncu_stream_t stream) {
    ReduceOps o;
    if (!lookup_arg(kind, t, &o)) { set_error("no arg-reduction kernel for kind %d type %d", kind, t); return PNCU_ERR_UNSUPPORTED; }
    int rc = ensure_init();
    if (rc) return rc;
    if (!out_index || count <= 0) { set_error("bad arguments"); return PNCU_ERR_ARGUMENT; }
    return o.finish(partials, count, out_index, NULL, as_stream(stream));
}

// ---- cross-GPU exchange fused into pass 2 (include/pnumpy_cuda.h) ---------------------------------------------
static int fill_peers(const pncu_peer_table* peers, int self, int64_t slot_base, PeerOut* po) {
    if (!peers || peers->n < 1 || peers->n > PNCU_MAX_PEERS || self < 0 || self >= peers->n) {
        set_error("peer table: n must be 1..%d and 0 <= self < n", PNCU_MAX_PEERS);
        return PNCU_ERR_ARGUMENT;
    }
    po->n = peers->n;
    po->self = self;
    po->slot_base = slot_base;
    for (int d = 0; d < peers->n; ++d) {
        if (!peers->slots[d] || !peers->arrived[d]) { set_error("peer table: null pointer for participant %d", d); return PNCU_ERR_ARGUMENT; }
        po->slots[d] = peers->slots[d];
        po->arrived[d] = peers->arrived[d];
    }
    return 0;
}

extern "C" int pncu_exchange_slices(void) { return EXCH_SLICES; }

extern "C" int pncu_reduce_exchange_finish(int func, int t, const void* local_partials, int64_t local_count,
                                           int64_t slot_base, const pncu_peer_table* peers, int self, int64_t total_count,
                                           void* out, const void* startVal, unsigned long long expected,
                                           unsigned long long* status, pncu_stream_t stream) {
    ReduceOps o;
    if (!lookup_reduce(func, t, &o)) { set_error("no reduction kernel for func %d type %d", func, t); return PNCU_ERR_UNSUPPORTED; }
    int rc = ensure_init();
    if (rc) return rc;
    if (!out || !local_partials || local_count <= 0 || total_count < local_count) { set_error("bad arguments"); return PNCU_ERR_ARGUMENT; }
    PeerOut po;
    if ((rc = fill_peers(peers, self, slot_base, &po))) return rc;
    return o.exchange_finish(local_partials, local_count, &po, total_count, out, startVal, expected, status, as_stream(stream));
}

extern "C" int pncu_argminmax_exchange_finish(int kind, int t, const void* local_partials, int64_t local_count,
                                              int64_t slot_base, const pncu_peer_table* peers, int self, int64_t total_count,
                                              int64_t* out_index, unsigned long long expected, unsigned long long* status,
                                              pncu_stream_t stream) {
    ReduceOps o;
    if (!lookup_arg(kind, t, &o)) { set_error("no arg-reduction kernel for kind %d type %d", kind, t); return PNCU_ERR_UNSUPPORTED; }
    int rc = ensure_init();
    if (rc) return rc;
    if (!out_index || !local_partials || local_count <= 0 || total_count < local_count) { set_error("bad arguments"); return PNCU_ERR_ARGUMENT; }
    PeerOut po;
    if ((rc = fill_peers(peers, self, slot_base, &po))) return rc;
    return o.exchange_finish(local_partials, local_count, &po, total_count, out_index, NULL, expected, status, as_stream(stream));
}

extern "C" int64_t pncu_last_tail_ns(void) {
    unsigned long long v = 0;
    if (ensure_init()) return -1;
    if (cudaMemcpyFromSymbol(&v, g_tail_ns, sizeof(v)) != cudaSuccess) { cudaGetLastError(); return -1; }
    return (int64_t)v;
}

extern "C" int64_t pncu_exchange_timeouts(void) {
    unsigned long long v = 0;
    if (ensure_init()) return -1;
    if (cudaMemcpyFromSymbol(&v, g_exchange_timeouts, sizeof(v)) != cudaSuccess) { cudaGetLastError(); return -1; }
    return (int64_t)v;
}

// ---- one-launch reductions (include/pnumpy_cuda.h) ---------------------------------------------------------
extern "C" int pncu_reduce_fused(int func, int t, const void* in, int64_t len, const pncu_fused_exchange* x, void* out,
                                 const void* startVal, pncu_stream_t stream) {
    ReduceOps o;
    if (!lookup_reduce(func, t, &o)) { set_error("no reduction kernel for func %d type %d", func, t); return PNCU_ERR_UNSUPPORTED; }
    int rc = ensure_init();
    if (rc) return rc;
    return o.fused(in, len, pncu_itemsize(t), 0, x, out, startVal, as_stream(stream));
}

extern "C" int pncu_argminmax_fused(int kind, int t, const void* in, int64_t len, int64_t index_base,
                                    const pncu_fused_exchange* x, int64_t* out_index, pncu_stream_t stream) {
    ReduceOps o;
    if (!lookup_arg(kind, t, &o)) { set_error("no arg-reduction kernel for kind %d type %d", kind, t); return PNCU_ERR_UNSUPPORTED; }
    int rc = ensure_init();
    if (rc) return rc;
    return o.fused(in, len, pncu_itemsize(t), index_base, x, out_index, NULL, as_stream(stream));
}

extern "C" int pncu_muladd_sum_fused(int t, const void* in1, const void* in2, const void* in3, int64_t len,
                                     const pncu_fused_exchange* x, void* out, pncu_stream_t stream) {
    MulAddSumOps o;
    if (!lookup_muladd_sum(t, &o)) { set_error("no fused multiply-add-sum kernel for type %d", t); return PNCU_ERR_UNSUPPORTED; }
    int rc = ensure_init();
    if (rc) return rc;
    return o.fused(in1, in2, in3, len, x, out, as_stream(stream));
}

extern "C" int pncu_fused_reset(pncu_stream_t stream) {
    int rc = ensure_init();
    if (rc) return rc;
    return reset_ticket(as_stream(stream));
}

extern "C" pncu_three_reduce_fn pncu_GetMulAddSumFast(int t) {
    MulAddSumOps o;
    return lookup_muladd_sum(t, &o) ? o.full : NULL;
}

extern "C" int pncu_muladd_sum_partials(int t, const void* in1, const void* in2, const void* in3, int64_t len,
                                        void* partials, pncu_stream_t stream) {
    MulAddSumOps o;
    if (!lookup_muladd_sum(t, &o)) { set_error("no fused multiply-add-sum kernel for type %d", t); return PNCU_ERR_UNSUPPORTED; }
    int rc = ensure_init();
    if (rc) return rc;
    if (len <= 0) return 0;
    return o.partials(in1, in2, in3, len, partials, as_stream(stream), nullptr);
}
