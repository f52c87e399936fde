// nte -- the per-GPU stream and shard dispatcher (see include/pnumpy_nte.h).
//
// replaces: CMathWorker / stMATH_WORKER_ITEM / stWorkerRing / WorkerThreadFunction
// (src/atop/threads.h:165-1171, src/atop/threads.cpp:72-191) as driven by the four hook-layer
// dispatchers (src/pnumpy/_pnumpy.cpp:638-996, 1105-1118).
//
// Reference                                   | here
// --------------------------------------------+---------------------------------------------------
// 16384-element chunks dealt round-robin to   | [0, n) cut into `lanes` contiguous shards at
// per-core channels (threads.h:319-389)       | multiples of 65536 elements; lane l -> GPU l % G
// futex-woken worker threads (threads.cpp:164)| persistent host threads, one per extra lane, woken
//                                             | by a condition variable; the caller runs lane 0
//                                             | (threads.h:1143: "main thread works as core 0")
// loop body on the chunk, in cache            | per lane: pinned staging ring -> H2D stream ->
//                                             | compute stream (libpnumpy_cuda launcher) -> D2H
//                                             | stream, slabs pipelined through 3 ring slots
// partials[chunk] then the same reduce on the | block partials gathered on GPU 0 (peer copies over
// caller thread (_pnumpy.cpp:666-700)         | NVLink), ONE finish kernel in fixed order
// spin until BlocksCompleted (threads.h:1148) | join lanes; stream-synchronize; return
// re-entrant call -> unthreaded (:1064-1073)  | try_lock fails -> NTE_DECLINED (caller's old loop)
#include <string.h>
#include <atomic>
#include <condition_variable>
#include <map>
#include <mutex>
#include <thread>
#include <vector>
#include "../../../include/pnumpy_nte.h"
#include "../atop/common.cuh"

using namespace pncu;

namespace {

constexpr int kRing = 3;          // staging slots per lane
constexpr int kMaxLanes = 16;     // the reference's hard maximum: 15 workers + caller (threads.h:56-66)
constexpr int64_t kAlign = 65536; // == pncu_reduce_block_elems()

struct Stats {
    std::atomic<int64_t> calls{0}, declined{0}, launches{0}, h2d{0}, d2h{0}, staged{0}, lanes_last{0}, gpus_last{0};
};
Stats g_stats;

struct Lane {
    int device = -1;
    bool ready = false;
    cudaStream_t s_h2d = nullptr, s_k = nullptr, s_d2h = nullptr;
    cudaEvent_t ev_h2d[kRing], ev_k[kRing], ev_d2h[kRing];
    char* d_buf[kRing][4];   // device slabs: in1, in2, out, in3 (in3: fused chain only, allocated lazily)
    char* h_buf[kRing][4];   // pinned staging slabs (allocated lazily, pageable operands only)
    size_t slab_bytes = 0;
    char* d_scalar = nullptr;   // 2 x 16 B for broadcast scalars
    char* d_partials = nullptr; // reduction block partials of this lane's shard
    size_t partials_bytes = 0;
    char* h_small = nullptr;    // pinned scratch for scalars / results (64 B)
};

struct Job;
typedef void (*LaneFn)(Lane&, Job&, int lane_index);

struct Worker {
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    LaneFn fn = nullptr;
    Job* job = nullptr;
    int lane_index = 0;
    bool has_work = false, done = false, quit = false;
};

struct State {
    std::mutex api_mu;            // one dispatch at a time (re-entrant callers are declined)
    bool inited = false;
    int ndev = 0;
    int workers = 3;              // extra lanes beside the caller's (reference default futex wake count is
                                  // min(11, cores-1), threads.h:67,891-894; per-op MaxThreads caps it at 3 or 7)
    bool threading = true;
    // Below the threshold a PCIe round trip cannot beat the caller's own loop.  `min_elems` is the base: it applies
    // to streaming element-wise ops on PINNED operands; the effective value is x4 when an operand is pageable (the
    // staging memcpy), /2 for transcendental ops (NumPy's loop is several times slower per element there), x2 for reductions
    // (no D2H to overlap).  Measured crossovers against NumPy's single-threaded loops on float32
    // (tools/threshold_probe.py, tools/pn_benchmark.py): add 2^19 pinned / 2^21 pageable, sin 2^17-2^18 / 2^19-2^20
    // (depends on how hard the arguments are for NumPy's sin), sum 2^20 / 2^22.
    int64_t min_elems = 1 << 19;
    size_t slab_bytes = 32u << 20;  // tools/e2e_probe.py on B200: 32 MiB slabs give 75 GB/s of PCIe traffic vs 73 at 8 MiB (pinned), 43 vs 31 (pageable, 4 lanes)
    Lane lanes[kMaxLanes];
    Worker* pool[kMaxLanes] = {nullptr};
};
State g;

// ---- per-call job ---------------------------------------------------------------------------------
enum JobKind { JOB_BINARY, JOB_UNARY, JOB_REDUCE, JOB_ARG, JOB_MULADD, JOB_MULADD_SUM };
enum PtrKind { PK_PAGEABLE = 0, PK_PINNED = 1, PK_DEVICE = 2, PK_MANAGED = 3 };

struct Job {
    JobKind kind;
    pncu_any_two_fn fn2 = nullptr;
    pncu_unary_fn fn1 = nullptr;
    pncu_any_three_fn fn3 = nullptr;
    int func = 0, atype = 0, argkind = 0;
    int isz_in = 0, isz_out = 0;
    const char* in1 = nullptr;
    const char* in2 = nullptr;
    const char* in3 = nullptr;
    char* out = nullptr;
    int64_t len = 0;
    int64_t s1 = 0, s2 = 0;     // input byte strides: 0 (scalar), the item size (contiguous) or anything else (gathered)
    int64_t so = 0;             // output byte stride (0 = not set: contiguous)
    int op_class = 0;           // NTE_CLASS_STREAM / _HEAVY / _REDUCE: scales the size threshold
    int pk_in1 = 0, pk_in2 = 0, pk_in3 = 0, pk_out = 0;
    int nlanes = 1;
    int64_t shard[kMaxLanes + 1];  // element boundaries
    int rc[kMaxLanes];
    char err[kMaxLanes][256];
    // reductions
    int partial_bytes = 0;
    int64_t slot0[kMaxLanes + 1];  // first partial slot of each lane
};

int fail(Job& j, int lane, int rc) {
    j.rc[lane] = rc;
    snprintf(j.err[lane], sizeof(j.err[lane]), "%s", pncu_last_error());
    return rc;
}
#define LCUDA(call)                                                             \
    do {                                                                        \
        cudaError_t _e = (call);                                                \
        if (_e != cudaSuccess) { cuda_fail(_e, #call, __FILE__, __LINE__); fail(j, li, (int)_e); return; } \
    } while (0)

int classify(const void* p) {
    cudaPointerAttributes at;
    cudaError_t e = cudaPointerGetAttributes(&at, p);
    if (e != cudaSuccess) { cudaGetLastError(); return PK_PAGEABLE; }
    switch (at.type) {
        case cudaMemoryTypeHost: return PK_PINNED;
        case cudaMemoryTypeDevice: return PK_DEVICE;
        case cudaMemoryTypeManaged: return PK_MANAGED;
        default: return PK_PAGEABLE;
    }
}

// ---- strided host operands ------------------------------------------------------------------------------
// NumPy hands an inner loop whatever strides the view has (a[::2], a[:, 3], a[::-1]); the reference runs its
// generic strided scalar loop per chunk on its worker threads (src/atop/ops_binary.cpp:563-569).  Here the
// lane thread gathers the elements of a slab into its pinned staging slab (and scatters results back), so
// the GPU always sees contiguous slabs and every lane does its own share of the host-side strided traffic.
template <class T>
void gather_t(char* dst, const char* src, int64_t cnt, int64_t stride) {
    T* d = reinterpret_cast<T*>(dst);
    for (int64_t i = 0; i < cnt; ++i) d[i] = *reinterpret_cast<const T*>(src + i * stride);
}
template <class T>
void scatter_t(char* dst, const char* src, int64_t cnt, int64_t stride) {
    const T* s = reinterpret_cast<const T*>(src);
    for (int64_t i = 0; i < cnt; ++i) *reinterpret_cast<T*>(dst + i * stride) = s[i];
}
void gather(char* dst, const char* src, int64_t cnt, int64_t stride, int isz) {
    switch (isz) {
        case 1: gather_t<uint8_t>(dst, src, cnt, stride); break;
        case 2: gather_t<uint16_t>(dst, src, cnt, stride); break;
        case 4: gather_t<uint32_t>(dst, src, cnt, stride); break;
        default: gather_t<uint64_t>(dst, src, cnt, stride); break;
    }
}
void scatter(char* dst, const char* src, int64_t cnt, int64_t stride, int isz) {
    switch (isz) {
        case 1: scatter_t<uint8_t>(dst, src, cnt, stride); break;
        case 2: scatter_t<uint16_t>(dst, src, cnt, stride); break;
        case 4: scatter_t<uint32_t>(dst, src, cnt, stride); break;
        default: scatter_t<uint64_t>(dst, src, cnt, stride); break;
    }
}
// byte extent [lo, lo + bytes) of `len` items of `isz` bytes at byte stride `stride` from `p`
void extent_of(const char* p, int64_t len, int64_t stride, int isz, const char** lo, int64_t* bytes) {
    const int64_t span = (len - 1) * stride;
    if (span >= 0) { *lo = p; *bytes = span + isz; }
    else { *lo = p + span; *bytes = -span + isz; }
}
bool stride_ok(const void* p, int64_t stride, int isz) {
    return (isz == 1 || isz == 2 || isz == 4 || isz == 8) && stride % isz == 0 && (uintptr_t)p % isz == 0;
}

// ---- lane resources -----------------------------------------------------------------------------------
int ensure_lane(int li) {
    Lane& L = g.lanes[li];
    const int dev = li % g.ndev;
    if (L.ready && L.slab_bytes == g.slab_bytes) return 0;
    PNCU_CUDA(cudaSetDevice(dev));
    if (!L.ready) {
        L.device = dev;
        PNCU_CUDA(cudaStreamCreateWithFlags(&L.s_h2d, cudaStreamNonBlocking));
        PNCU_CUDA(cudaStreamCreateWithFlags(&L.s_k, cudaStreamNonBlocking));
        PNCU_CUDA(cudaStreamCreateWithFlags(&L.s_d2h, cudaStreamNonBlocking));
        for (int r = 0; r < kRing; ++r) {
            PNCU_CUDA(cudaEventCreateWithFlags(&L.ev_h2d[r], cudaEventDisableTiming));
            PNCU_CUDA(cudaEventCreateWithFlags(&L.ev_k[r], cudaEventDisableTiming));
            PNCU_CUDA(cudaEventCreateWithFlags(&L.ev_d2h[r], cudaEventDisableTiming));
            for (int k = 0; k < 4; ++k) { L.d_buf[r][k] = nullptr; L.h_buf[r][k] = nullptr; }
        }
        PNCU_CUDA(cudaMalloc((void**)&L.d_scalar, 64));
        PNCU_CUDA(cudaHostAlloc((void**)&L.h_small, 64, cudaHostAllocPortable));
    } else {
        for (int r = 0; r < kRing; ++r)
            for (int k = 0; k < 4; ++k) {
                if (L.d_buf[r][k]) cudaFree(L.d_buf[r][k]);
                if (L.h_buf[r][k]) cudaFreeHost(L.h_buf[r][k]);
                L.d_buf[r][k] = nullptr; L.h_buf[r][k] = nullptr;
            }
    }
    // device slabs now; pinned slabs lazily (only pageable operands need them)
    for (int r = 0; r < kRing; ++r)
        for (int k = 0; k < 3; ++k) PNCU_CUDA(cudaMalloc((void**)&L.d_buf[r][k], g.slab_bytes));
    L.slab_bytes = g.slab_bytes;
    L.ready = true;
    return 0;
}

int ensure_pinned_slab(Lane& L, int r, int k) {
    if (L.h_buf[r][k]) return 0;
    PNCU_CUDA(cudaHostAlloc((void**)&L.h_buf[r][k], L.slab_bytes, cudaHostAllocPortable));
    return 0;
}

int ensure_third_input(Lane& L) {
    for (int r = 0; r < kRing; ++r)
        if (!L.d_buf[r][3]) PNCU_CUDA(cudaMalloc((void**)&L.d_buf[r][3], L.slab_bytes));
    return 0;
}

int ensure_partials(Lane& L, size_t bytes) {
    if (L.partials_bytes >= bytes) return 0;
    if (L.d_partials) cudaFree(L.d_partials);
    L.d_partials = nullptr;
    L.partials_bytes = 0;
    size_t want = bytes < (1u << 20) ? (1u << 20) : bytes * 2;
    PNCU_CUDA(cudaMalloc((void**)&L.d_partials, want));
    L.partials_bytes = want;
    return 0;
}

int64_t slab_elems_for(size_t slab_bytes, int64_t wide, int64_t shard_elems, bool staged);

// ---- one lane of an element-wise job: slabs through the H2D -> kernel -> D2H pipeline --------------------
void lane_elementwise(Lane& L, Job& j, int li) {
    if (cudaSetDevice(L.device) != cudaSuccess) { cudaGetLastError(); fail(j, li, PNCU_ERR_NO_DEVICE); return; }
    const int64_t e0 = j.shard[li], e1 = j.shard[li + 1];
    if (e0 >= e1) return;
    const bool binary = j.kind == JOB_BINARY;
    const int64_t wide = j.isz_in > j.isz_out ? j.isz_in : j.isz_out;
    const bool vec1 = j.s1 != 0, vec2 = binary && j.s2 != 0;
    const bool out_strided = j.so != j.isz_out;
    const bool stage_out = j.pk_out == PK_PAGEABLE || out_strided;
    const bool stage_in = (vec1 && (j.pk_in1 == PK_PAGEABLE || j.s1 != j.isz_in)) || (vec2 && (j.pk_in2 == PK_PAGEABLE || j.s2 != j.isz_in));
    const int64_t slab_elems = slab_elems_for(L.slab_bytes, wide, e1 - e0, stage_in || stage_out);
    const int64_t nslabs = (e1 - e0 + slab_elems - 1) / slab_elems;

    // broadcast scalars: one tiny copy per call
    const char* d_s1 = nullptr;
    const char* d_s2 = nullptr;
    if (!vec1) {
        memcpy(L.h_small, j.in1, j.isz_in);
        LCUDA(cudaMemcpyAsync(L.d_scalar, L.h_small, j.isz_in, cudaMemcpyHostToDevice, L.s_k));
        d_s1 = L.d_scalar;
    }
    if (binary && !vec2) {
        memcpy(L.h_small + 16, j.in2, j.isz_in);
        LCUDA(cudaMemcpyAsync(L.d_scalar + 16, L.h_small + 16, j.isz_in, cudaMemcpyHostToDevice, L.s_k));
        d_s2 = L.d_scalar + 16;
    }

    auto slab_range = [&](int64_t i, int64_t* off, int64_t* cnt) {
        *off = e0 + i * slab_elems;
        *cnt = (e1 - *off) < slab_elems ? (e1 - *off) : slab_elems;
    };
    auto retire = [&](int64_t i) -> int {  // slab i's result is on the host once ev_d2h fires
        const int r = (int)(i % kRing);
        cudaError_t e = cudaEventSynchronize(L.ev_d2h[r]);
        if (e != cudaSuccess) return cuda_fail(e, "cudaEventSynchronize(d2h)", __FILE__, __LINE__);
        if (stage_out) {
            int64_t off, cnt;
            slab_range(i, &off, &cnt);
            if (out_strided) scatter(j.out + off * j.so, L.h_buf[r][2], cnt, j.so, j.isz_out);
            else memcpy(j.out + off * j.isz_out, L.h_buf[r][2], (size_t)(cnt * j.isz_out));
            g_stats.staged += cnt * j.isz_out;
        }
        return 0;
    };

    for (int64_t i = 0; i < nslabs; ++i) {
        const int r = (int)(i % kRing);
        int rc;
        if (i >= kRing && (rc = retire(i - kRing))) { fail(j, li, rc); return; }
        int64_t off, cnt;
        slab_range(i, &off, &cnt);
        const size_t in_bytes = (size_t)(cnt * j.isz_in), out_bytes = (size_t)(cnt * j.isz_out);
        // ---- inputs -> device slab
        for (int k = 0; k < 2; ++k) {
            if (k == 0 ? !vec1 : !vec2) continue;
            const int64_t stride = k == 0 ? j.s1 : j.s2;
            const char* src = (k == 0 ? j.in1 : j.in2) + off * stride;
            const int pk = k == 0 ? j.pk_in1 : j.pk_in2;
            if (pk == PK_PAGEABLE || stride != j.isz_in) {
                if ((rc = ensure_pinned_slab(L, r, k))) { fail(j, li, rc); return; }
                if (stride != j.isz_in) gather(L.h_buf[r][k], src, cnt, stride, j.isz_in);
                else memcpy(L.h_buf[r][k], src, in_bytes);
                g_stats.staged += in_bytes;
                src = L.h_buf[r][k];
            }
            LCUDA(cudaMemcpyAsync(L.d_buf[r][k], src, in_bytes, cudaMemcpyHostToDevice, L.s_h2d));
            g_stats.h2d += in_bytes;
        }
        LCUDA(cudaEventRecord(L.ev_h2d[r], L.s_h2d));
        LCUDA(cudaStreamWaitEvent(L.s_k, L.ev_h2d[r], 0));
        // ---- kernel
        if (binary)
            rc = j.fn2(vec1 ? L.d_buf[r][0] : d_s1, vec2 ? L.d_buf[r][1] : d_s2, L.d_buf[r][2], cnt,
                       vec1 ? j.isz_in : 0, vec2 ? j.isz_in : 0, j.isz_out, (pncu_stream_t)L.s_k);
        else
            rc = j.fn1(L.d_buf[r][0], L.d_buf[r][2], cnt, j.isz_in, j.isz_out, (pncu_stream_t)L.s_k);
        if (rc) { fail(j, li, rc); return; }
        g_stats.launches += 1;
        LCUDA(cudaEventRecord(L.ev_k[r], L.s_k));
        LCUDA(cudaStreamWaitEvent(L.s_d2h, L.ev_k[r], 0));
        // the next H2D into this slot must not overtake this kernel's reads
        LCUDA(cudaStreamWaitEvent(L.s_h2d, L.ev_k[r], 0));
        // ---- result -> host
        char* dst = j.out + off * j.isz_out;
        if (stage_out) {
            if ((rc = ensure_pinned_slab(L, r, 2))) { fail(j, li, rc); return; }
            dst = L.h_buf[r][2];
        }
        LCUDA(cudaMemcpyAsync(dst, L.d_buf[r][2], out_bytes, cudaMemcpyDeviceToHost, L.s_d2h));
        g_stats.d2h += out_bytes;
        LCUDA(cudaEventRecord(L.ev_d2h[r], L.s_d2h));
        // the next kernel writing this slot's out slab must wait for this D2H
        LCUDA(cudaStreamWaitEvent(L.s_k, L.ev_d2h[r], 0));
    }
    for (int64_t i = nslabs > kRing ? nslabs - kRing : 0; i < nslabs; ++i) {
        int rc = retire(i);
        if (rc) { fail(j, li, rc); return; }
    }
}

// ---- one lane of a reduction / arg-reduction: slabs H2D -> pass-1 partials on this lane's GPU ------------
void lane_reduce(Lane& L, Job& j, int li) {
    if (cudaSetDevice(L.device) != cudaSuccess) { cudaGetLastError(); fail(j, li, PNCU_ERR_NO_DEVICE); return; }
    const int64_t e0 = j.shard[li], e1 = j.shard[li + 1];
    if (e0 >= e1) return;
    const int64_t nslots = j.slot0[li + 1] - j.slot0[li];
    int rc;
    if ((rc = ensure_partials(L, (size_t)nslots * j.partial_bytes))) { fail(j, li, rc); return; }
    const int64_t slab_elems = slab_elems_for(L.slab_bytes, j.isz_in, e1 - e0, j.pk_in1 == PK_PAGEABLE || j.s1 != j.isz_in);
    const int64_t nslabs = (e1 - e0 + slab_elems - 1) / slab_elems;
    for (int64_t i = 0; i < nslabs; ++i) {
        const int r = (int)(i % kRing);
        const int64_t off = e0 + i * slab_elems;
        const int64_t cnt = (e1 - off) < slab_elems ? (e1 - off) : slab_elems;
        const size_t in_bytes = (size_t)(cnt * j.isz_in);
        const char* src = j.in1 + off * j.s1;
        if (i >= kRing) {  // slot reuse: the kernel that read it must be done before we overwrite staging
            cudaError_t e = cudaEventSynchronize(L.ev_k[r]);
            if (e != cudaSuccess) { cuda_fail(e, "cudaEventSynchronize(k)", __FILE__, __LINE__); fail(j, li, (int)e); return; }
        }
        if (j.pk_in1 == PK_PAGEABLE || j.s1 != j.isz_in) {
            if ((rc = ensure_pinned_slab(L, r, 0))) { fail(j, li, rc); return; }
            if (j.s1 != j.isz_in) gather(L.h_buf[r][0], src, cnt, j.s1, j.isz_in);
            else memcpy(L.h_buf[r][0], src, in_bytes);
            g_stats.staged += in_bytes;
            src = L.h_buf[r][0];
        }
        LCUDA(cudaMemcpyAsync(L.d_buf[r][0], src, in_bytes, cudaMemcpyHostToDevice, L.s_h2d));
        g_stats.h2d += in_bytes;
        LCUDA(cudaEventRecord(L.ev_h2d[r], L.s_h2d));
        LCUDA(cudaStreamWaitEvent(L.s_k, L.ev_h2d[r], 0));
        const int64_t slot = pncu_reduce_partial_count(j.atype, off - e0);
        char* parts = L.d_partials + slot * j.partial_bytes;
        if (j.kind == JOB_REDUCE)
            rc = pncu_reduce_partials(j.func, j.atype, L.d_buf[r][0], cnt, parts, (pncu_stream_t)L.s_k);
        else
            rc = pncu_argminmax_partials(j.argkind, j.atype, L.d_buf[r][0], cnt, off, parts, (pncu_stream_t)L.s_k);
        if (rc) { fail(j, li, rc); return; }
        g_stats.launches += 1;
        LCUDA(cudaEventRecord(L.ev_k[r], L.s_k));
        LCUDA(cudaStreamWaitEvent(L.s_h2d, L.ev_k[r], 0));
    }
    cudaError_t e = cudaStreamSynchronize(L.s_k);
    if (e != cudaSuccess) { cuda_fail(e, "cudaStreamSynchronize", __FILE__, __LINE__); fail(j, li, (int)e); }
}

// ---- one lane of the fused chain: three input slabs -> a*b+c (-> D2H) or -> block partials of its sum ------
void lane_fused(Lane& L, Job& j, int li) {
    if (cudaSetDevice(L.device) != cudaSuccess) { cudaGetLastError(); fail(j, li, PNCU_ERR_NO_DEVICE); return; }
    const int64_t e0 = j.shard[li], e1 = j.shard[li + 1];
    if (e0 >= e1) return;
    const bool to_sum = j.kind == JOB_MULADD_SUM;
    const int isz = j.isz_in;
    int rc;
    if ((rc = ensure_third_input(L))) { fail(j, li, rc); return; }
    if (to_sum && (rc = ensure_partials(L, (size_t)(j.slot0[li + 1] - j.slot0[li]) * j.partial_bytes))) { fail(j, li, rc); return; }
    const int64_t slab_elems = slab_elems_for(L.slab_bytes, isz, e1 - e0,
                                              j.pk_in1 == PK_PAGEABLE || j.pk_in2 == PK_PAGEABLE || j.pk_in3 == PK_PAGEABLE ||
                                                  (!to_sum && j.pk_out == PK_PAGEABLE));
    const int64_t nslabs = (e1 - e0 + slab_elems - 1) / slab_elems;
    const char* ins[3] = {j.in1, j.in2, j.in3};
    const int pks[3] = {j.pk_in1, j.pk_in2, j.pk_in3};
    const int slot_of[3] = {0, 1, 3};  // d_buf / h_buf index of each input (2 is `out`)

    auto slab_range = [&](int64_t i, int64_t* off, int64_t* cnt) {
        *off = e0 + i * slab_elems;
        *cnt = (e1 - *off) < slab_elems ? (e1 - *off) : slab_elems;
    };
    // slot reuse: a sum only needs the kernel that read the slot to be done; an element-wise job needs
    // the slot's result on the host (and copies it out of staging for a pageable `out`)
    auto retire = [&](int64_t i) -> int {
        const int r = (int)(i % kRing);
        cudaError_t e = cudaEventSynchronize(to_sum ? L.ev_k[r] : L.ev_d2h[r]);
        if (e != cudaSuccess) return cuda_fail(e, "cudaEventSynchronize(fused)", __FILE__, __LINE__);
        if (!to_sum && j.pk_out == PK_PAGEABLE) {
            int64_t off, cnt;
            slab_range(i, &off, &cnt);
            memcpy(j.out + off * isz, L.h_buf[r][2], (size_t)(cnt * isz));
            g_stats.staged += cnt * isz;
        }
        return 0;
    };

    for (int64_t i = 0; i < nslabs; ++i) {
        const int r = (int)(i % kRing);
        if (i >= kRing && (rc = retire(i - kRing))) { fail(j, li, rc); return; }
        int64_t off, cnt;
        slab_range(i, &off, &cnt);
        const size_t bytes = (size_t)(cnt * isz);
        for (int k = 0; k < 3; ++k) {
            const char* src = ins[k] + off * isz;
            if (pks[k] == PK_PAGEABLE) {
                if ((rc = ensure_pinned_slab(L, r, slot_of[k]))) { fail(j, li, rc); return; }
                memcpy(L.h_buf[r][slot_of[k]], src, bytes);
                g_stats.staged += bytes;
                src = L.h_buf[r][slot_of[k]];
            }
            LCUDA(cudaMemcpyAsync(L.d_buf[r][slot_of[k]], src, bytes, cudaMemcpyHostToDevice, L.s_h2d));
            g_stats.h2d += bytes;
        }
        LCUDA(cudaEventRecord(L.ev_h2d[r], L.s_h2d));
        LCUDA(cudaStreamWaitEvent(L.s_k, L.ev_h2d[r], 0));
        if (to_sum) {
            const int64_t slot = pncu_reduce_partial_count(j.atype, off - e0);
            rc = pncu_muladd_sum_partials(j.atype, L.d_buf[r][0], L.d_buf[r][1], L.d_buf[r][3], cnt,
                                          L.d_partials + slot * j.partial_bytes, (pncu_stream_t)L.s_k);
        } else {
            rc = j.fn3(L.d_buf[r][0], L.d_buf[r][1], L.d_buf[r][3], L.d_buf[r][2], cnt, (pncu_stream_t)L.s_k);
        }
        if (rc) { fail(j, li, rc); return; }
        g_stats.launches += 1;
        LCUDA(cudaEventRecord(L.ev_k[r], L.s_k));
        LCUDA(cudaStreamWaitEvent(L.s_h2d, L.ev_k[r], 0));
        if (!to_sum) {
            LCUDA(cudaStreamWaitEvent(L.s_d2h, L.ev_k[r], 0));
            char* dst = j.out + off * isz;
            if (j.pk_out == PK_PAGEABLE) {
                if ((rc = ensure_pinned_slab(L, r, 2))) { fail(j, li, rc); return; }
                dst = L.h_buf[r][2];
            }
            LCUDA(cudaMemcpyAsync(dst, L.d_buf[r][2], bytes, cudaMemcpyDeviceToHost, L.s_d2h));
            g_stats.d2h += bytes;
            LCUDA(cudaEventRecord(L.ev_d2h[r], L.s_d2h));
            LCUDA(cudaStreamWaitEvent(L.s_k, L.ev_d2h[r], 0));
        }
    }
    if (to_sum) {
        cudaError_t e = cudaStreamSynchronize(L.s_k);
        if (e != cudaSuccess) { cuda_fail(e, "cudaStreamSynchronize", __FILE__, __LINE__); fail(j, li, (int)e); }
    } else {
        for (int64_t i = nslabs > kRing ? nslabs - kRing : 0; i < nslabs; ++i)
            if ((rc = retire(i))) { fail(j, li, rc); return; }
    }
}

// ---- worker pool -----------------------------------------------------------------------------------------
void worker_main(Worker* w) {
    std::unique_lock<std::mutex> lk(w->mu);
    for (;;) {
        w->cv.wait(lk, [&] { return w->has_work || w->quit; });
        if (w->quit) return;
        w->has_work = false;
        lk.unlock();
        w->fn(g.lanes[w->lane_index], *w->job, w->lane_index);
        lk.lock();
        w->done = true;
        w->cv.notify_all();
    }
}

void run_lanes(Job& j, LaneFn fn) {
    for (int li = 1; li < j.nlanes; ++li) {
        Worker*& w = g.pool[li];
        if (!w) {
            w = new Worker();
            w->th = std::thread(worker_main, w);
        }
        std::lock_guard<std::mutex> lk(w->mu);
        w->fn = fn; w->job = &j; w->lane_index = li; w->done = false; w->has_work = true;
        w->cv.notify_all();
    }
    fn(g.lanes[0], j, 0);  // the caller is lane 0
    for (int li = 1; li < j.nlanes; ++li) {
        Worker* w = g.pool[li];
        std::unique_lock<std::mutex> lk(w->mu);
        w->cv.wait(lk, [&] { return w->done; });
    }
}

// the size below which the caller's own loop is faster (see State::min_elems).  The base is stated in 4-byte
// elements and applied to BYTES of the widest operand: NumPy's loops run at a byte rate, not an element rate (an int8
// add of 2^20 elements is 1 MiB per operand and over in 30 us).  1-byte types get another x3: their NumPy loops stay
// in cache and vectorise 32 lanes wide (pn.benchmark() on the B200 box: int8 a+b breaks even near 6 MiB per operand,
// int16 / float32 near 2 MiB).
int64_t threshold_for(int op_class, bool any_pageable, int wide_itemsize = 4) {
    int64_t bytes = g.min_elems * 4;
    if (any_pageable) bytes *= 4;
    if (op_class == NTE_CLASS_HEAVY) bytes /= 2;
    if (op_class == NTE_CLASS_REDUCE) bytes *= 2;
    if (wide_itemsize == 1) bytes *= 3;
    if (wide_itemsize < 1) wide_itemsize = 1;
    return bytes / wide_itemsize;   // in elements of the widest operand
}
// slab length for a shard of `shard_elems`: big shards use the full slab; when an operand has to be STAGED (pageable
// or strided: a host-side copy per slab) smaller shards are cut in up to four slabs of at least 2 MiB so that the
// staging copy of slab i+1 overlaps the DMA and kernel of slab i (pageable sum at 2^24 float32: 2.9 -> 1.9 ms)
int64_t slab_elems_for(size_t slab_bytes, int64_t wide, int64_t shard_elems, bool staged) {
    const int64_t full = ((int64_t)(slab_bytes / wide) / kAlign) * kAlign;
    if (!staged) return full;  // pinned operands are DMA'd in place: fewer, larger transfers win (measured)
    int64_t quarter = (((shard_elems + 3) / 4 + kAlign - 1) / kAlign) * kAlign;
    int64_t floor_e = ((((int64_t)2 << 20) / wide + kAlign - 1) / kAlign) * kAlign;
    int64_t e = quarter > floor_e ? quarter : floor_e;
    return e < full ? e : full;
}

// lanes for a call: min(op_max_workers, workers) + 1, never more than one per 65536-element unit
int plan(Job& j, int max_workers) {
    int lanes = 1;
    if (g.threading) {
        int extra = max_workers < g.workers ? max_workers : g.workers;
        if (extra < 0) extra = 0;
        lanes = extra + 1;
    }
    if (lanes > kMaxLanes) lanes = kMaxLanes;
    const int64_t units = (j.len + kAlign - 1) / kAlign;
    if (lanes > units) lanes = (int)units;
    if (lanes < 1) lanes = 1;
    j.nlanes = lanes;
    const int64_t per = (units + lanes - 1) / lanes;
    for (int l = 0; l <= lanes; ++l) {
        int64_t e = (int64_t)l * per * kAlign;
        j.shard[l] = e < j.len ? e : j.len;
    }
    for (int l = 0; l < lanes; ++l) { j.rc[l] = 0; j.err[l][0] = 0; }
    int gpus = lanes < g.ndev ? lanes : g.ndev;
    g_stats.lanes_last = lanes;
    g_stats.gpus_last = gpus;
    for (int l = 0; l < lanes; ++l) {
        int rc = ensure_lane(l);
        if (rc) return rc;
    }
    return 0;
}

int first_error(Job& j) {
    for (int l = 0; l < j.nlanes; ++l)
        if (j.rc[l]) { set_error("lane %d (GPU %d): %s", l, g.lanes[l].device, j.err[l]); return j.rc[l]; }
    return 0;
}

int decline() { g_stats.declined += 1; return NTE_DECLINED; }

bool host_kind(int pk) { return pk == PK_PAGEABLE || pk == PK_PINNED; }

// operands that already live in device / managed memory: run in place on that GPU, no PCIe
// the GPU a resident pointer belongs to (managed memory: wherever it was last prefetched, else GPU 0)
int device_of(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) == cudaSuccess && at.device >= 0 && at.device < g.ndev) return at.device;
    cudaGetLastError();
    return 0;
}

// managed operands: make the pages resident on the GPU before the kernel touches them, otherwise the
// kernel demand-faults them over PCIe at a fraction of the copy rate.  A no-op once they are there.
// Prefetching a range that is already resident still costs ~1 ms of driver time per GiB (page-table
// walk), 5x the kernel it precedes.  So a range is prefetched the FIRST time it is seen on a GPU and
// trusted afterwards; if host code touched some pages in between, the kernel demand-faults those back
// (correct, slower, and then they are resident again).  Guarded by g.api_mu (held by every dispatch).
std::map<const void*, size_t> g_prefetched[kMaxLanes];
std::mutex g_prefetch_mu;  // the lanes of a sharded resident call prefetch concurrently
std::map<const void*, size_t> g_managed_allocs;  // blocks handed out by nte_alloc_managed

int prefetch_managed(const void* p, int pk, size_t bytes, Lane& L) {
    if (pk != PK_MANAGED || !bytes) return 0;
    std::lock_guard<std::mutex> lk(g_prefetch_mu);
    std::map<const void*, size_t>& seen = g_prefetched[L.device];
    auto it = seen.find(p);
    if (it != seen.end() && it->second >= bytes) return 0;
    PNCU_CUDA(cudaMemPrefetchAsync(p, bytes, L.device, L.s_k));
    if (seen.size() > 4096) seen.clear();
    seen[p] = bytes;
    // the pages leave whichever GPU had them: drop the other devices' records that overlap the range
    for (int d = 0; d < kMaxLanes; ++d) {
        if (d == L.device) continue;
        std::map<const void*, size_t>& other = g_prefetched[d];
        for (auto o = other.begin(); o != other.end();) {
            const char* lo = (const char*)o->first;
            if (lo < (const char*)p + bytes && (const char*)p < lo + o->second) o = other.erase(o);
            else ++o;
        }
    }
    return 0;
}

void forget_prefetched(const void* lo, size_t bytes) {
    std::lock_guard<std::mutex> lk(g_prefetch_mu);
    for (int d = 0; d < kMaxLanes; ++d) {
        std::map<const void*, size_t>& seen = g_prefetched[d];
        for (auto it = seen.lower_bound(lo); it != seen.end() && (const char*)it->first < (const char*)lo + bytes;) it = seen.erase(it);
    }
}

int run_resident(Job& j) {
    const int dev = device_of(j.out);
    int rc = ensure_lane(dev);
    if (rc) return rc;
    g_stats.lanes_last = 1;
    g_stats.gpus_last = 1;
    Lane& L = g.lanes[dev];
    PNCU_CUDA(cudaSetDevice(L.device));
    const bool binary = j.kind == JOB_BINARY;
    const char* in1 = j.in1;
    const char* in2 = j.in2;
    // a host scalar (NumPy's `a + 5.0`) next to resident arrays: one 8-byte copy
    if (j.s1 == 0 && host_kind(j.pk_in1)) {
        memcpy(L.h_small, j.in1, j.isz_in);
        PNCU_CUDA(cudaMemcpyAsync(L.d_scalar, L.h_small, j.isz_in, cudaMemcpyHostToDevice, L.s_k));
        in1 = L.d_scalar;
    }
    if (binary && j.s2 == 0 && host_kind(j.pk_in2)) {
        memcpy(L.h_small + 16, j.in2, j.isz_in);
        PNCU_CUDA(cudaMemcpyAsync(L.d_scalar + 16, L.h_small + 16, j.isz_in, cudaMemcpyHostToDevice, L.s_k));
        in2 = L.d_scalar + 16;
    }
    const char* lo;
    int64_t bytes;
    if (j.s1) {
        extent_of(j.in1, j.len, j.s1, j.isz_in, &lo, &bytes);
        if ((rc = prefetch_managed(lo, j.pk_in1, (size_t)bytes, L))) return rc;
    }
    if (binary && j.s2) {
        extent_of(j.in2, j.len, j.s2, j.isz_in, &lo, &bytes);
        if ((rc = prefetch_managed(lo, j.pk_in2, (size_t)bytes, L))) return rc;
    }
    extent_of(j.out, j.len, j.so, j.isz_out, &lo, &bytes);
    if ((rc = prefetch_managed(lo, j.pk_out, (size_t)bytes, L))) return rc;
    if (binary)
        rc = j.fn2(in1, in2, j.out, j.len, j.s1, j.s2, j.so, (pncu_stream_t)L.s_k);
    else
        rc = j.fn1(in1, j.out, j.len, j.s1, j.so, (pncu_stream_t)L.s_k);
    if (rc) return rc;
    g_stats.launches += 1;
    PNCU_CUDA(cudaStreamSynchronize(L.s_k));
    return 0;
}

// ---- managed operands spread over all GPUs ----------------------------------------------------------------
// A managed (cudaMallocManaged) ndarray is addressable from every GPU, so a large one does not have to live
// on one: shard l of every operand -- the same contiguous split at multiples of 65 536 elements the host path
// uses -- is prefetched to GPU l the first time it is seen there, and every call then runs one kernel per GPU on
// its own shard, in place, with no PCIe and no exchange (reductions: block partials gathered on GPU 0 as in the
// host path).  np.add(a, b, out=c) on managed arrays thus runs at G times one GPU's HBM rate.
constexpr size_t kShardResidentMinBytes = (size_t)64 << 20;  // smaller arrays stay on one GPU

bool shard_resident(const Job& j, bool binary) {
    if (g.ndev < 2 || !g.threading) return false;
    if (j.pk_out != PK_MANAGED || j.so != j.isz_out) return false;
    if (j.s1 != 0 && (j.pk_in1 != PK_MANAGED || j.s1 != j.isz_in)) return false;
    if (binary && j.s2 != 0 && (j.pk_in2 != PK_MANAGED || j.s2 != j.isz_in)) return false;
    return (size_t)j.len * (size_t)(j.isz_in > j.isz_out ? j.isz_in : j.isz_out) >= kShardResidentMinBytes;
}

int plan_one_lane_per_gpu(Job& j) {
    const int64_t units = (j.len + kAlign - 1) / kAlign;
    int lanes = g.ndev < units ? g.ndev : (int)units;
    if (lanes < 1) lanes = 1;
    j.nlanes = lanes;
    const int64_t per = (units + lanes - 1) / lanes;
    for (int l = 0; l <= lanes; ++l) {
        const int64_t e = (int64_t)l * per * kAlign;
        j.shard[l] = e < j.len ? e : j.len;
    }
    for (int l = 0; l < lanes; ++l) {
        j.rc[l] = 0; j.err[l][0] = 0;
        int rc = ensure_lane(l);
        if (rc) return rc;
    }
    g_stats.lanes_last = lanes;
    g_stats.gpus_last = lanes;
    return 0;
}

void lane_resident(Lane& L, Job& j, int li) {
    if (cudaSetDevice(L.device) != cudaSuccess) { cudaGetLastError(); fail(j, li, PNCU_ERR_NO_DEVICE); return; }
    const int64_t e0 = j.shard[li], e1 = j.shard[li + 1];
    if (e0 >= e1) return;
    const int64_t n = e1 - e0;
    const bool binary = j.kind == JOB_BINARY;
    const char* in1 = j.s1 ? j.in1 + e0 * j.isz_in : j.in1;
    const char* in2 = binary ? (j.s2 ? j.in2 + e0 * j.isz_in : j.in2) : nullptr;
    char* out = j.out + e0 * j.isz_out;
    int rc;
    if (j.s1 == 0 && host_kind(j.pk_in1)) {
        memcpy(L.h_small, j.in1, j.isz_in);
        LCUDA(cudaMemcpyAsync(L.d_scalar, L.h_small, j.isz_in, cudaMemcpyHostToDevice, L.s_k));
        in1 = L.d_scalar;
    }
    if (binary && j.s2 == 0 && host_kind(j.pk_in2)) {
        memcpy(L.h_small + 16, j.in2, j.isz_in);
        LCUDA(cudaMemcpyAsync(L.d_scalar + 16, L.h_small + 16, j.isz_in, cudaMemcpyHostToDevice, L.s_k));
        in2 = L.d_scalar + 16;
    }
    if (j.s1 && (rc = prefetch_managed(in1, j.pk_in1, (size_t)n * j.isz_in, L))) { fail(j, li, rc); return; }
    if (binary && j.s2 && (rc = prefetch_managed(in2, j.pk_in2, (size_t)n * j.isz_in, L))) { fail(j, li, rc); return; }
    if ((rc = prefetch_managed(out, j.pk_out, (size_t)n * j.isz_out, L))) { fail(j, li, rc); return; }
    if (binary) rc = j.fn2(in1, in2, out, n, j.s1, j.s2, j.so, (pncu_stream_t)L.s_k);
    else rc = j.fn1(in1, out, n, j.s1, j.so, (pncu_stream_t)L.s_k);
    if (rc) { fail(j, li, rc); return; }
    g_stats.launches += 1;
}

void lane_reduce_resident(Lane& L, Job& j, int li) {
    if (cudaSetDevice(L.device) != cudaSuccess) { cudaGetLastError(); fail(j, li, PNCU_ERR_NO_DEVICE); return; }
    const int64_t e0 = j.shard[li], e1 = j.shard[li + 1];
    if (e0 >= e1) return;
    const int64_t n = e1 - e0;
    const char* in = j.in1 + e0 * j.isz_in;
    int rc;
    if ((rc = ensure_partials(L, (size_t)(j.slot0[li + 1] - j.slot0[li]) * j.partial_bytes))) { fail(j, li, rc); return; }
    if ((rc = prefetch_managed(in, j.pk_in1, (size_t)n * j.isz_in, L))) { fail(j, li, rc); return; }
    if (j.kind == JOB_REDUCE) rc = pncu_reduce_partials(j.func, j.atype, in, n, L.d_partials, (pncu_stream_t)L.s_k);
    else rc = pncu_argminmax_partials(j.argkind, j.atype, in, n, e0, L.d_partials, (pncu_stream_t)L.s_k);
    if (rc) { fail(j, li, rc); return; }
    g_stats.launches += 1;
}

// fused chain on sharded managed operands: a*b+c in place per GPU, or pass 1 of its sum into the lane's partials
void lane_fused_resident(Lane& L, Job& j, int li) {
    if (cudaSetDevice(L.device) != cudaSuccess) { cudaGetLastError(); fail(j, li, PNCU_ERR_NO_DEVICE); return; }
    const int64_t e0 = j.shard[li], e1 = j.shard[li + 1];
    if (e0 >= e1) return;
    const int64_t n = e1 - e0;
    const size_t bytes = (size_t)n * j.isz_in;
    const bool to_sum = j.kind == JOB_MULADD_SUM;
    const char* a = j.in1 + e0 * j.isz_in;
    const char* b = j.in2 + e0 * j.isz_in;
    const char* c = j.in3 + e0 * j.isz_in;
    int rc;
    if ((rc = prefetch_managed(a, j.pk_in1, bytes, L)) || (rc = prefetch_managed(b, j.pk_in2, bytes, L)) ||
        (rc = prefetch_managed(c, j.pk_in3, bytes, L))) { fail(j, li, rc); return; }
    if (!to_sum) {
        char* o = j.out + e0 * j.isz_in;
        if ((rc = prefetch_managed(o, j.pk_out, bytes, L))) { fail(j, li, rc); return; }
        rc = j.fn3(a, b, c, o, n, (pncu_stream_t)L.s_k);
    } else {
        if ((rc = ensure_partials(L, (size_t)(j.slot0[li + 1] - j.slot0[li]) * j.partial_bytes))) { fail(j, li, rc); return; }
        rc = pncu_muladd_sum_partials(j.atype, a, b, c, n, L.d_partials, (pncu_stream_t)L.s_k);
    }
    if (rc) { fail(j, li, rc); return; }
    g_stats.launches += 1;
}

// Resident shards need no host-side copying, so no worker threads either: the calling thread queues the work of
// every GPU (asynchronous launches, ~5 us each) and then waits for all of them -- a condition-variable wake-up per
// lane would cost more than a 128 MiB shard takes to stream.
void run_resident_lanes(Job& j, LaneFn fn) {
    for (int li = 0; li < j.nlanes; ++li) fn(g.lanes[li], j, li);
    for (int li = 0; li < j.nlanes; ++li) {
        if (j.rc[li]) continue;
        Lane& L = g.lanes[li];
        cudaError_t e = cudaSetDevice(L.device);
        if (e == cudaSuccess) e = cudaStreamSynchronize(L.s_k);
        if (e != cudaSuccess) { cuda_fail(e, "cudaStreamSynchronize(resident shard)", __FILE__, __LINE__); fail(j, li, (int)e); }
    }
}

int run_resident_sharded(Job& j) {
    int rc = plan_one_lane_per_gpu(j);
    if (rc) return rc;
    run_resident_lanes(j, lane_resident);
    return first_error(j);
}

}  // namespace

// =======================================================================================================
extern "C" {

int nte_init(void) {
    std::lock_guard<std::mutex> lk(g.api_mu);
    if (g.inited) return 0;
    int n = 0;
    int rc = pncu_init(&n);
    if (rc) return rc;
    g.ndev = n > kMaxLanes ? kMaxLanes : n;
    const char* env = getenv("PNUMPY_CUDA_MIN_ELEMS");
    if (env && *env) g.min_elems = atoll(env);
    env = getenv("PNUMPY_CUDA_SLAB_BYTES");
    if (env && *env) g.slab_bytes = (size_t)atoll(env);
    env = getenv("PNUMPY_CUDA_DEVICES");  // use only the first k devices
    if (env && *env && atoi(env) > 0 && atoi(env) < g.ndev) g.ndev = atoi(env);
    // peer access so that the reduction gather can go GPU -> GPU 0 over NVLink
    for (int a = 0; a < g.ndev; ++a)
        for (int b = 0; b < g.ndev; ++b) {
            if (a == b) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, a, b) == cudaSuccess && can) {
                cudaSetDevice(a);
                cudaError_t e = cudaDeviceEnablePeerAccess(b, 0);
                if (e != cudaSuccess) cudaGetLastError();
            } else {
                cudaGetLastError();
            }
        }
    cudaSetDevice(0);
    g.inited = true;
    return 0;
}

int nte_shutdown(void) {
    std::lock_guard<std::mutex> lk(g.api_mu);
    for (int l = 1; l < kMaxLanes; ++l) {
        Worker* w = g.pool[l];
        if (!w) continue;
        { std::lock_guard<std::mutex> wl(w->mu); w->quit = true; w->cv.notify_all(); }
        w->th.join();
        delete w;
        g.pool[l] = nullptr;
    }
    for (int l = 0; l < kMaxLanes; ++l) {
        Lane& L = g.lanes[l];
        if (!L.ready) continue;
        cudaSetDevice(L.device);
        for (int r = 0; r < kRing; ++r) {
            for (int k = 0; k < 4; ++k) {
                if (L.d_buf[r][k]) cudaFree(L.d_buf[r][k]);
                if (L.h_buf[r][k]) cudaFreeHost(L.h_buf[r][k]);
            }
            cudaEventDestroy(L.ev_h2d[r]); cudaEventDestroy(L.ev_k[r]); cudaEventDestroy(L.ev_d2h[r]);
        }
        if (L.d_scalar) cudaFree(L.d_scalar);
        if (L.d_partials) cudaFree(L.d_partials);
        if (L.h_small) cudaFreeHost(L.h_small);
        cudaStreamDestroy(L.s_h2d); cudaStreamDestroy(L.s_k); cudaStreamDestroy(L.s_d2h);
        L = Lane();
    }
    g.inited = false;
    return 0;
}

int nte_device_count(void) { return g.inited ? g.ndev : 0; }

int nte_max_workers(void) { return kMaxLanes - 1; }
int nte_set_workers(int workers) {
    // CMathWorker::SetFutexWakeup, src/atop/threads.h:852-872: clamp, return the previous value
    int prev = g.workers;
    if (workers < 1) workers = 1;
    if (workers > nte_max_workers()) workers = nte_max_workers();
    g.workers = workers;
    return prev;
}
int nte_get_workers(void) { return g.workers; }
void nte_set_threading(int enabled) { g.threading = enabled != 0; }
int nte_get_threading(void) { return g.threading ? 1 : 0; }
int64_t nte_set_min_elems(int64_t n) { int64_t p = g.min_elems; g.min_elems = n < 0 ? 0 : n; return p; }
int64_t nte_get_min_elems(void) { return g.min_elems; }
int64_t nte_set_slab_bytes(int64_t bytes) {
    std::lock_guard<std::mutex> lk(g.api_mu);
    int64_t p = (int64_t)g.slab_bytes;
    const int64_t unit = kAlign * 8;
    if (bytes < unit) bytes = unit;
    g.slab_bytes = (size_t)((bytes / unit) * unit);
    return p;
}

void* nte_alloc_pinned(size_t bytes) {
    if (pncu_init(nullptr)) return nullptr;
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
void nte_free_pinned(void* ptr) {
    if (ptr && cudaFreeHost(ptr) != cudaSuccess) cudaGetLastError();
}

void* nte_alloc_managed(size_t bytes) {
    if (pncu_init(nullptr)) return nullptr;
    void* p = nullptr;
    if (cudaMallocManaged(&p, bytes ? bytes : 1, cudaMemAttachGlobal) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    std::lock_guard<std::mutex> lk(g.api_mu);
    g_managed_allocs[p] = bytes ? bytes : 1;
    return p;
}
void nte_free_managed(void* ptr) {
    if (!ptr) return;
    {
        std::lock_guard<std::mutex> lk(g.api_mu);
        auto it = g_managed_allocs.find(ptr);
        const size_t bytes = it != g_managed_allocs.end() ? it->second : 1;
        if (it != g_managed_allocs.end()) g_managed_allocs.erase(it);
        forget_prefetched(ptr, bytes);  // views into the block are remembered by their own base address
    }
    if (cudaFree(ptr) != cudaSuccess) cudaGetLastError();
}

void nte_get_stats(nte_stats* o) {
    if (!o) return;
    o->calls = g_stats.calls; o->declined = g_stats.declined; o->launches = g_stats.launches;
    o->h2d_bytes = g_stats.h2d; o->d2h_bytes = g_stats.d2h; o->staged_bytes = g_stats.staged;
    o->lanes_last = g_stats.lanes_last; o->gpus_last = g_stats.gpus_last;
}
void nte_reset_stats(void) {
    g_stats.calls = 0; g_stats.declined = 0; g_stats.launches = 0; g_stats.h2d = 0; g_stats.d2h = 0;
    g_stats.staged = 0; g_stats.lanes_last = 0; g_stats.gpus_last = 0;
}

static int dispatch_elementwise(Job& j, int max_workers) {
    if (!g.inited) { set_error("nte_init() has not succeeded"); return PNCU_ERR_NO_DEVICE; }
    if (j.len <= 0 || j.len < threshold_for(NTE_CLASS_HEAVY, false, 8)) return decline();  // below every class's threshold
    std::unique_lock<std::mutex> lk(g.api_mu, std::try_to_lock);
    if (!lk.owns_lock()) return decline();  // re-entrant / concurrent caller: run the old loop
    const bool binary = j.kind == JOB_BINARY;
    j.pk_in1 = classify(j.in1);
    j.pk_in2 = binary ? classify(j.in2) : PK_PAGEABLE;
    j.pk_out = classify(j.out);
    const bool any_resident = !host_kind(j.pk_out) || (j.s1 != 0 && !host_kind(j.pk_in1)) ||
                              (binary && j.s2 != 0 && !host_kind(j.pk_in2));
    if (any_resident) {
        // every ARRAY operand must be resident (a host array next to a managed one goes to the caller's
        // loop, which may read managed memory); host scalars are copied by run_resident
        const bool all_ok = !host_kind(j.pk_out) && (j.s1 == 0 || !host_kind(j.pk_in1)) &&
                            (!binary || j.s2 == 0 || !host_kind(j.pk_in2));
        if (!all_ok) return decline();
        int rc = shard_resident(j, binary) ? run_resident_sharded(j) : run_resident(j);
        if (!rc) g_stats.calls += 1;
        return rc;
    }
    // an op whose output does not depend on its input (isnan / isinf / isfinite of integers): shipping the input
    // over PCIe to memset the output would be absurd -- NumPy's own loop is a memset already
    if (j.op_class == NTE_CLASS_CONST) return decline();
    {
        const bool pageable = j.pk_out == PK_PAGEABLE || (j.s1 != 0 && j.pk_in1 == PK_PAGEABLE) ||
                              (binary && j.s2 != 0 && j.pk_in2 == PK_PAGEABLE);
        if (j.len < threshold_for(j.op_class, pageable, j.isz_in > j.isz_out ? j.isz_in : j.isz_out)) return decline();
    }
    // host operands: contiguous ones are DMA'd (pinned) or memcpy'd (pageable) into the slabs, any
    // other stride (a multiple of the item size, negative allowed) is gathered / scattered by the lane
    if (j.s1 == 0 && (!binary || j.s2 == 0)) return decline();
    if (!stride_ok(j.in1, j.s1, j.isz_in) || (binary && !stride_ok(j.in2, j.s2, j.isz_in)) || !stride_ok(j.out, j.so, j.isz_out))
        return decline();
    // exact aliasing of out with an input is fine (every element is staged before its result is written
    // back); any other overlap needs the sequential semantics of the caller's loop
    {
        const char* olo;
        int64_t ob;
        extent_of(j.out, j.len, j.so, j.isz_out, &olo, &ob);
        const char* ins[2] = {j.in1, binary ? j.in2 : nullptr};
        const int64_t strides[2] = {j.s1, j.s2};
        for (int k = 0; k < 2; ++k) {
            if (!ins[k]) continue;
            const char* ilo;
            int64_t ib;
            extent_of(ins[k], strides[k] ? j.len : 1, strides[k], j.isz_in, &ilo, &ib);
            const bool overlap = ilo < olo + ob && olo < ilo + ib;
            if (overlap && !(ins[k] == j.out && strides[k] == j.so && j.isz_in == j.isz_out)) return decline();
        }
    }
    int rc = plan(j, max_workers);
    if (rc) return rc;
    run_lanes(j, lane_elementwise);
    rc = first_error(j);
    if (!rc) g_stats.calls += 1;
    return rc;
}

int nte_binary(pncu_any_two_fn launcher, int in_itemsize, int out_itemsize, const void* in1,
               const void* in2, void* out, int64_t len, int64_t s1, int64_t s2, int64_t so, int max_workers) {
    if (!launcher) return decline();
    if (so == 0) return decline();
    Job j;
    j.kind = JOB_BINARY; j.fn2 = launcher; j.isz_in = in_itemsize; j.isz_out = out_itemsize;
    j.in1 = (const char*)in1; j.in2 = (const char*)in2; j.out = (char*)out; j.len = len; j.s1 = s1; j.s2 = s2; j.so = so;
    return dispatch_elementwise(j, max_workers);
}

int nte_unary(pncu_unary_fn launcher, int in_itemsize, int out_itemsize, const void* in, void* out,
              int64_t len, int64_t s1, int64_t so, int max_workers) {
    return nte_unary_class(launcher, in_itemsize, out_itemsize, in, out, len, s1, so, max_workers, NTE_CLASS_STREAM);
}

int nte_unary_class(pncu_unary_fn launcher, int in_itemsize, int out_itemsize, const void* in, void* out,
                    int64_t len, int64_t s1, int64_t so, int max_workers, int op_class) {
    if (!launcher) return decline();
    if (so == 0 || s1 == 0) return decline();
    Job j;
    j.kind = JOB_UNARY; j.fn1 = launcher; j.isz_in = in_itemsize; j.isz_out = out_itemsize;
    j.in1 = (const char*)in; j.out = (char*)out; j.len = len; j.s1 = s1; j.so = so; j.op_class = op_class;
    return dispatch_elementwise(j, max_workers);
}

// gather every lane's partials on GPU 0, fold there in fixed order, bring the scalar back
static int finish_on_gpu0(Job& j, void* host_out, int out_bytes, const void* host_start) {
    Lane& L0 = g.lanes[0];
    PNCU_CUDA(cudaSetDevice(L0.device));
    const int64_t total = j.slot0[j.nlanes];
    char* all = L0.d_partials;
    if (j.nlanes > 1) {
        // lane 0's buffer holds its own slots at [0, n0); append the others behind them
        if ((size_t)total * j.partial_bytes > L0.partials_bytes) {
            // grow, preserving lane 0's partials
            char* bigger = nullptr;
            PNCU_CUDA(cudaMalloc((void**)&bigger, (size_t)total * j.partial_bytes * 2));
            PNCU_CUDA(cudaMemcpyAsync(bigger, L0.d_partials, (size_t)j.slot0[1] * j.partial_bytes, cudaMemcpyDeviceToDevice, L0.s_k));
            PNCU_CUDA(cudaStreamSynchronize(L0.s_k));
            cudaFree(L0.d_partials);
            L0.d_partials = bigger;
            L0.partials_bytes = (size_t)total * j.partial_bytes * 2;
            all = bigger;
        }
        for (int l = 1; l < j.nlanes; ++l) {
            const int64_t cnt = j.slot0[l + 1] - j.slot0[l];
            if (cnt <= 0) continue;
            Lane& L = g.lanes[l];
            // peer copy (NVLink P2P when enabled, else staged by the driver); same-GPU lanes: plain D2D
            PNCU_CUDA(cudaMemcpyPeerAsync(all + j.slot0[l] * j.partial_bytes, L0.device, L.d_partials, L.device,
                                          (size_t)cnt * j.partial_bytes, L0.s_k));
        }
    }
    char* d_res = L0.d_scalar + 32;
    int rc;
    if (j.kind == JOB_REDUCE) {
        const void* d_start = nullptr;
        if (host_start) {
            memcpy(L0.h_small, host_start, out_bytes);
            PNCU_CUDA(cudaMemcpyAsync(L0.d_scalar, L0.h_small, out_bytes, cudaMemcpyHostToDevice, L0.s_k));
            d_start = L0.d_scalar;
        }
        rc = pncu_reduce_finish(j.func, j.atype, all, total, d_res, d_start, (pncu_stream_t)L0.s_k);
    } else {
        rc = pncu_argminmax_finish(j.argkind, j.atype, all, total, (int64_t*)d_res, (pncu_stream_t)L0.s_k);
    }
    if (rc) return rc;
    g_stats.launches += 1;
    PNCU_CUDA(cudaMemcpyAsync(L0.h_small + 32, d_res, out_bytes, cudaMemcpyDeviceToHost, L0.s_k));
    PNCU_CUDA(cudaStreamSynchronize(L0.s_k));
    memcpy(host_out, L0.h_small + 32, out_bytes);
    g_stats.d2h += out_bytes;
    return 0;
}

static int dispatch_reduce(Job& j, void* host_out, int out_bytes, const void* host_start, int max_workers) {
    if (!g.inited) { set_error("nte_init() has not succeeded"); return PNCU_ERR_NO_DEVICE; }
    if (j.len <= 0 || j.len < threshold_for(NTE_CLASS_HEAVY, false, 8)) return decline();
    std::unique_lock<std::mutex> lk(g.api_mu, std::try_to_lock);
    if (!lk.owns_lock()) return decline();
    j.pk_in1 = classify(j.in1);
    if (host_kind(j.pk_in1) && j.len < threshold_for(NTE_CLASS_REDUCE, j.pk_in1 == PK_PAGEABLE, j.isz_in)) return decline();
    int rc;
    if (!host_kind(j.pk_in1)) {
        // resident input: both passes on its GPU (contiguous only: pass 1 streams whole blocks)
        if (j.s1 != j.isz_in) return decline();
        if (g.ndev > 1 && g.threading && j.pk_in1 == PK_MANAGED && (size_t)j.len * j.isz_in >= kShardResidentMinBytes) {
            // a large managed array: one shard per GPU (see run_resident_sharded), partials gathered on GPU 0
            if ((rc = plan_one_lane_per_gpu(j))) return rc;
            for (int l = 0; l <= j.nlanes; ++l) j.slot0[l] = pncu_reduce_partial_count(j.atype, j.shard[l]);
            run_resident_lanes(j, lane_reduce_resident);
            if ((rc = first_error(j))) return rc;
            rc = finish_on_gpu0(j, host_out, out_bytes, host_start);
            if (!rc) g_stats.calls += 1;
            return rc;
        }
        const int dev = device_of(j.in1);
        if ((rc = ensure_lane(dev))) return rc;
        g_stats.lanes_last = 1;
        g_stats.gpus_last = 1;
        Lane& L = g.lanes[dev];
        PNCU_CUDA(cudaSetDevice(L.device));
        if ((rc = prefetch_managed(j.in1, j.pk_in1, (size_t)j.len * j.isz_in, L))) return rc;
        const int64_t slots = pncu_reduce_partial_count(j.atype, j.len);
        if ((rc = ensure_partials(L, (size_t)slots * j.partial_bytes))) return rc;
        if (j.kind == JOB_REDUCE) rc = pncu_reduce_partials(j.func, j.atype, j.in1, j.len, L.d_partials, (pncu_stream_t)L.s_k);
        else rc = pncu_argminmax_partials(j.argkind, j.atype, j.in1, j.len, 0, L.d_partials, (pncu_stream_t)L.s_k);
        if (rc) return rc;
        g_stats.launches += 1;
        j.nlanes = 1; j.slot0[0] = 0; j.slot0[1] = slots;
        // finish_on_gpu0 works on lanes[0]; alias it when the data lives elsewhere
        if (&L != &g.lanes[0]) {
            PNCU_CUDA(cudaStreamSynchronize(L.s_k));
            if ((rc = ensure_lane(0))) return rc;
            if ((rc = ensure_partials(g.lanes[0], (size_t)slots * j.partial_bytes))) return rc;
            PNCU_CUDA(cudaMemcpyPeer(g.lanes[0].d_partials, g.lanes[0].device, L.d_partials, L.device, (size_t)slots * j.partial_bytes));
        }
        rc = finish_on_gpu0(j, host_out, out_bytes, host_start);
        if (!rc) g_stats.calls += 1;
        return rc;
    }
    if ((rc = plan(j, max_workers))) return rc;
    for (int l = 0; l <= j.nlanes; ++l) j.slot0[l] = pncu_reduce_partial_count(j.atype, j.shard[l]);
    run_lanes(j, lane_reduce);
    if ((rc = first_error(j))) return rc;
    rc = finish_on_gpu0(j, host_out, out_bytes, host_start);
    if (!rc) g_stats.calls += 1;
    return rc;
}

int nte_reduce(int func, int atype, const void* in, void* out, int use_out_as_start, int64_t len,
               int64_t strideIn, int max_workers) {
    const int isz = pncu_itemsize(atype);
    const int pb = pncu_reduce_partial_bytes(func, atype);
    if (!isz || !pb || strideIn == 0 || !stride_ok(in, strideIn, isz)) return decline();
    Job j;
    j.kind = JOB_REDUCE; j.func = func; j.atype = atype; j.isz_in = isz; j.isz_out = isz; j.partial_bytes = pb;
    j.in1 = (const char*)in; j.len = len; j.s1 = strideIn;
    return dispatch_reduce(j, out, isz, use_out_as_start ? out : nullptr, max_workers);
}

int nte_argminmax(int kind, int atype, const void* in, int64_t len, int64_t* out_index, int max_workers) {
    const int isz = pncu_itemsize(atype);
    if (!isz || !pncu_GetArgMinMaxOpFast(kind, atype)) return decline();
    Job j;
    j.kind = JOB_ARG; j.argkind = kind; j.atype = atype; j.isz_in = isz; j.isz_out = 8; j.partial_bytes = 16;
    j.in1 = (const char*)in; j.len = len; j.s1 = isz;
    return dispatch_reduce(j, out_index, 8, nullptr, max_workers);
}

// fused chain: element-wise (out != NULL) or summed (host_out)
static int dispatch_fused(Job& j, void* host_out, int max_workers) {
    if (!g.inited) { set_error("nte_init() has not succeeded"); return PNCU_ERR_NO_DEVICE; }
    if (j.len <= 0 || j.len < threshold_for(NTE_CLASS_HEAVY, false, 8)) return decline();
    std::unique_lock<std::mutex> lk(g.api_mu, std::try_to_lock);
    if (!lk.owns_lock()) return decline();
    const bool to_sum = j.kind == JOB_MULADD_SUM;
    j.pk_in1 = classify(j.in1); j.pk_in2 = classify(j.in2); j.pk_in3 = classify(j.in3);
    j.pk_out = to_sum ? j.pk_in1 : classify(j.out);
    const int pks[4] = {j.pk_in1, j.pk_in2, j.pk_in3, j.pk_out};
    int nhost = 0, npageable = 0;
    for (int k = 0; k < 4; ++k) { nhost += host_kind(pks[k]); npageable += pks[k] == PK_PAGEABLE; }
    if (nhost != 0 && nhost != 4) return decline();  // host and resident arrays mixed: the caller's loops
    if (nhost == 4 && j.len < threshold_for(NTE_CLASS_STREAM, npageable != 0, j.isz_in)) return decline();
    const size_t bytes = (size_t)j.len * j.isz_in;
    int rc;
    if (nhost == 0) {
        // resident operands: in place on their GPU, no PCIe
        const bool all_managed = j.pk_in1 == PK_MANAGED && j.pk_in2 == PK_MANAGED && j.pk_in3 == PK_MANAGED &&
                                 (to_sum || j.pk_out == PK_MANAGED);
        if (g.ndev > 1 && g.threading && all_managed && bytes >= kShardResidentMinBytes) {
            // large managed operands live sharded over the GPUs (run_resident_sharded): so does the fused chain
            if ((rc = plan_one_lane_per_gpu(j))) return rc;
            if (to_sum)
                for (int l = 0; l <= j.nlanes; ++l) j.slot0[l] = pncu_reduce_partial_count(j.atype, j.shard[l]);
            run_resident_lanes(j, lane_fused_resident);
            if ((rc = first_error(j))) return rc;
            if (to_sum) {
                j.kind = JOB_REDUCE;
                j.func = PNCU_ADD;
                if ((rc = finish_on_gpu0(j, host_out, j.isz_in, nullptr))) return rc;
            }
            g_stats.calls += 1;
            return 0;
        }
        const int dev = device_of(j.in1);
        if ((rc = ensure_lane(dev))) return rc;
        g_stats.lanes_last = 1;
        g_stats.gpus_last = 1;
        Lane& L = g.lanes[dev];
        PNCU_CUDA(cudaSetDevice(L.device));
        if ((rc = prefetch_managed(j.in1, j.pk_in1, bytes, L))) return rc;
        if ((rc = prefetch_managed(j.in2, j.pk_in2, bytes, L))) return rc;
        if ((rc = prefetch_managed(j.in3, j.pk_in3, bytes, L))) return rc;
        if (!to_sum) {
            if ((rc = prefetch_managed(j.out, j.pk_out, bytes, L))) return rc;
            if ((rc = j.fn3(j.in1, j.in2, j.in3, j.out, j.len, (pncu_stream_t)L.s_k))) return rc;
            g_stats.launches += 1;
            PNCU_CUDA(cudaStreamSynchronize(L.s_k));
        } else {
            pncu_three_reduce_fn full = pncu_GetMulAddSumFast(j.atype);
            char* d_res = L.d_scalar + 32;
            if ((rc = full(j.in1, j.in2, j.in3, d_res, j.len, (pncu_stream_t)L.s_k))) return rc;
            g_stats.launches += 2;
            PNCU_CUDA(cudaMemcpyAsync(L.h_small + 32, d_res, j.isz_in, cudaMemcpyDeviceToHost, L.s_k));
            PNCU_CUDA(cudaStreamSynchronize(L.s_k));
            memcpy(host_out, L.h_small + 32, j.isz_in);
            g_stats.d2h += j.isz_in;
        }
        g_stats.calls += 1;
        return 0;
    }
    if (!to_sum) {
        // `out` may alias an input exactly; partial overlap is declined
        const char* ins[3] = {j.in1, j.in2, j.in3};
        for (int k = 0; k < 3; ++k) {
            const bool overlap = ins[k] < j.out + bytes && j.out < ins[k] + bytes;
            if (overlap && ins[k] != j.out) return decline();
        }
    }
    if ((rc = plan(j, max_workers))) return rc;
    if (to_sum)
        for (int l = 0; l <= j.nlanes; ++l) j.slot0[l] = pncu_reduce_partial_count(j.atype, j.shard[l]);
    run_lanes(j, lane_fused);
    if ((rc = first_error(j))) return rc;
    if (to_sum) {
        j.kind = JOB_REDUCE;  // the partials are those of the ADD reduction: ordinary gather + finish
        j.func = PNCU_ADD;
        rc = finish_on_gpu0(j, host_out, j.isz_in, nullptr);
        if (rc) return rc;
    }
    g_stats.calls += 1;
    return 0;
}

int nte_muladd(int atype, const void* in1, const void* in2, const void* in3, void* out, int64_t len, int max_workers) {
    pncu_any_three_fn fn = pncu_GetMulAddFast(atype);
    if (!fn) return decline();
    Job j;
    j.kind = JOB_MULADD; j.fn3 = fn; j.atype = atype; j.isz_in = j.isz_out = pncu_itemsize(atype);
    j.in1 = (const char*)in1; j.in2 = (const char*)in2; j.in3 = (const char*)in3; j.out = (char*)out; j.len = len;
    j.s1 = j.s2 = j.isz_in;
    return dispatch_fused(j, nullptr, max_workers);
}

int nte_muladd_sum(int atype, const void* in1, const void* in2, const void* in3, void* out, int64_t len, int max_workers) {
    if (!pncu_GetMulAddSumFast(atype)) return decline();
    Job j;
    j.kind = JOB_MULADD_SUM; j.atype = atype; j.isz_in = j.isz_out = pncu_itemsize(atype);
    j.partial_bytes = pncu_reduce_partial_bytes(PNCU_ADD, atype);
    j.in1 = (const char*)in1; j.in2 = (const char*)in2; j.in3 = (const char*)in3; j.len = len;
    j.s1 = j.s2 = j.isz_in;
    return dispatch_fused(j, out, max_workers);
}

}  // extern "C"
