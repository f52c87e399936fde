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
// futex-woken worker threads (threads.cpp:164)| persistent host threads, one per extra lane, that spin
//                                             | for a while after a job and then sleep in a futex
//                                             | (pool.h); the caller runs lane 0 (threads.h:1143:
//                                             | "main thread works as core 0")
// loop body on the chunk, in cache            | host operands: pinned staging ring -> H2D stream ->
//                                             | compute stream (libpnumpy_cuda launcher) -> D2H
//                                             | stream, slabs pipelined through 3 ring slots;
//                                             | resident operands (managed / device memory): ONE
//                                             | launch per GPU, in place, no PCIe
// partials[chunk] then the same reduce on the | resident: ONE kernel per GPU does pass 1, pushes the
// caller thread (_pnumpy.cpp:666-700)         | shard's block partials into GPU 0's table over NVLink
//                                             | and GPU 0's last CTA folds the table and writes the
//                                             | result into mapped host memory (pncu_*_fused);
//                                             | host operands: partials gathered on GPU 0 + finish
// spin until BlocksCompleted (threads.h:1148) | wait for the workers / the result flag; return
// re-entrant call -> unthreaded (:1064-1073)  | try_lock fails -> NTE_DECLINED (caller's old loop)
#include <pthread.h>
#include <string.h>
#include <atomic>
#include <mutex>
#include <vector>
#include "../../../include/pnumpy_nte.h"
#include "../atop/common.cuh"
#include "pool.h"

using namespace pncu;

namespace {

constexpr int kRing = 3;          // staging slots per lane
constexpr int kMaxLanes = 16;     // the reference's hard maximum: 15 workers + caller (threads.h:56-66)
constexpr int64_t kAlign = 65536; // == pncu_reduce_block_elems()
constexpr int64_t kBlockBytes = 65536;  // bytes of input per reduction slot (checked against pncu_reduce_partial_count in nte_init)

// optional per-call trace (nte_set_trace): monotonic ns at the stations of the last dispatch
enum { TR_ENTER, TR_LOCKED, TR_PLANNED, TR_PLACED, TR_POSTED, TR_LAUNCHED, TR_JOINED, TR_DONE, TR_N };
int64_t g_trace[TR_N];
bool g_trace_on = false;
#define TRACE(i) do { if (g_trace_on) g_trace[i] = (int64_t)nte::now_ns(); } while (0)

struct Stats {
    std::atomic<int64_t> calls{0}, declined{0}, launches{0}, h2d{0}, d2h{0}, staged{0}, lanes_last{0}, gpus_last{0},
        prefetches{0};
};
Stats g_stats;

// layout of a lane's 256-byte mapped host scratch (`h_small`) and device scratch (`d_scalar`)
constexpr int kOffScalar1 = 0, kOffScalar2 = 16, kOffResult = 32, kOffStart = 48, kOffFlag = 64;
constexpr int kOffArrived = 128;  // d_scalar of lane 0 only: arrival counter of the reduction exchange

struct Lane {
    int device = -1;
    bool ready = false;
    cudaStream_t s_h2d = nullptr, s_k = nullptr, s_d2h = nullptr;
    cudaEvent_t ev_h2d[kRing], ev_k[kRing], ev_d2h[kRing];
    char* d_buf[kRing][4];   // device slabs: in1, in2, out, in3 (in3: fused chain only, allocated lazily)
    char* h_buf[kRing][4];   // pinned staging slabs (allocated lazily, pageable operands only)
    bool slabs = false;      // device slabs allocated (host operands only need them)
    size_t slab_bytes = 0;
    char* d_scalar = nullptr;   // 256 B of device scratch (see kOff*)
    char* d_partials = nullptr; // reduction block partials of this lane's shard (host-operand path)
    size_t partials_bytes = 0;
    char* h_small = nullptr;    // 256 B of MAPPED pinned scratch: scalars in, results and the completion flag out
    uint64_t seq = 0;           // completion values handed to one-launch reductions
};

struct Job;
typedef void (*LaneFn)(Lane&, Job&, int lane_index);

struct State {
    std::mutex api_mu;            // one dispatch at a time (re-entrant callers are declined)
    bool inited = false;
    std::atomic<bool> forked{false};
    int ndev = 0;
    bool peer_all = false;        // every GPU can store into GPU 0's memory (the reduction exchange needs it)
    int workers = 3;              // extra lanes beside the caller's (reference default futex wake count is
                                  // min(11, cores-1), threads.h:67,891-894; per-op MaxThreads caps it at 3 or 7)
    bool threading = true;
    // Below the threshold a PCIe round trip cannot beat the caller's own loop.  `min_elems` is the base: it applies
    // to streaming element-wise ops on PINNED operands; the effective value is x4 when an operand is pageable (the
    // staging memcpy), /2 for transcendental ops (NumPy's loop is several times slower per element there), x2 for reductions
    // (no D2H to overlap).  Measured crossovers against NumPy's single-threaded loops on float32
    // (tools/threshold_probe.py, tools/pn_benchmark.py): add 2^19 pinned / 2^21 pageable, sin 2^17-2^18 / 2^19-2^20
    // (depends on how hard the arguments are for NumPy's sin), sum 2^20 / 2^22.
    // RESIDENT operands (managed / device memory) never see a threshold: they always run on the GPU.
    int64_t min_elems = 1 << 19;
    size_t slab_bytes = 32u << 20;  // tools/e2e_probe.py on B200: 32 MiB slabs give 75 GB/s of PCIe traffic vs 73 at 8 MiB (pinned), 43 vs 31 (pageable, 4 lanes)
    uint64_t spin_ns = 200000;      // how long a worker spins for the next job before it sleeps
    int64_t fault_countdown = 0;    // fault injection (tests): the n-th GPU dispatch from now fails
    Lane lanes[kMaxLanes];
    nte::Worker* pool[kMaxLanes] = {nullptr};
    // reduction exchange of the resident multi-GPU path: GPU 0 is the root
    char* x_table = nullptr;        // slot table on GPU 0
    size_t x_table_bytes = 0;
    unsigned long long x_expected = 0;
};
State g;

// ---- per-call job ---------------------------------------------------------------------------------
enum JobKind { JOB_BINARY, JOB_UNARY, JOB_REDUCE, JOB_ARG, JOB_MULADD, JOB_MULADD_SUM };

struct Operand {
    int kind = MK_UNKNOWN;
    int device = -1;
    const char* base = nullptr;     // start of the registered allocation (managed blocks: its head page is special)
    std::vector<ResidencyNote>* notes = nullptr;
    bool resident() const { return kind == MK_DEVICE || kind == MK_MANAGED; }
    bool host() const { return kind == MK_PAGEABLE || kind == MK_PINNED; }
};

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
    int64_t s1 = 0, s2 = 0;     // input byte strides: 0 (scalar), the item size (contiguous) or anything else
    int64_t so = 0;             // output byte stride
    int op_class = 0;           // NTE_CLASS_STREAM / _HEAVY / _REDUCE: scales the size threshold
    Operand k1, k2, k3, ko;
    int nlanes = 1;
    int lane_id[kMaxLanes];        // physical lane (-> GPU lane_id % G) of logical lane l
    int64_t shard[kMaxLanes + 1];  // element boundaries
    int rc[kMaxLanes];
    char err[kMaxLanes][256];
    // reductions
    int partial_bytes = 0;
    int64_t slot0[kMaxLanes + 1];  // first partial slot of each lane
    // one-launch reductions
    void* host_out = nullptr;
    const void* host_start = nullptr;
    int out_bytes = 0;
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

// the calling thread's current device is left as it was found (torch / cupy code in the same thread)
struct DeviceGuard {
    int prev = -1;
    DeviceGuard() { if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; } }
    ~DeviceGuard() { if (prev >= 0 && cudaSetDevice(prev) != cudaSuccess) cudaGetLastError(); }
};

// ---- what memory does an operand live in? ------------------------------------------------------------------
bool lookup_registry(const void* p, Operand* o) {
    AllocInfo ai;
    if (!registry_find(p, &ai)) return false;
    o->kind = ai.kind;
    o->device = ai.device;
    o->base = ai.base;
    o->notes = ai.notes;
    return true;
}
void classify_driver(const void* p, Operand* o) {
    cudaPointerAttributes at;
    cudaError_t e = cudaPointerGetAttributes(&at, p);
    if (e != cudaSuccess) { cudaGetLastError(); o->kind = MK_PAGEABLE; return; }
    switch (at.type) {
        case cudaMemoryTypeHost: o->kind = MK_PINNED; break;
        case cudaMemoryTypeDevice: o->kind = MK_DEVICE; o->device = at.device; break;
        case cudaMemoryTypeManaged: o->kind = MK_MANAGED; break;
        default: o->kind = MK_PAGEABLE; break;
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
// streams, events and the small scratch blocks: everything a RESIDENT call needs
int ensure_lane(int li) {
    Lane& L = g.lanes[li];
    if (L.ready) return 0;
    const int dev = li % g.ndev;
    PNCU_CUDA(cudaSetDevice(dev));
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
    PNCU_CUDA(cudaMalloc((void**)&L.d_scalar, 256));
    PNCU_CUDA(cudaMemset(L.d_scalar, 0, 256));
    PNCU_CUDA(cudaHostAlloc((void**)&L.h_small, 256, cudaHostAllocPortable | cudaHostAllocMapped));
    memset(L.h_small, 0, 256);
    L.ready = true;
    return 0;
}

// + the device slabs of the staging ring: host operands only
int ensure_slabs(int li) {
    int rc = ensure_lane(li);
    if (rc) return rc;
    Lane& L = g.lanes[li];
    if (L.slabs && L.slab_bytes == g.slab_bytes) return 0;
    PNCU_CUDA(cudaSetDevice(L.device));
    for (int r = 0; r < kRing; ++r)
        for (int k = 0; k < 4; ++k) {
            if (L.d_buf[r][k]) cudaFree(L.d_buf[r][k]);
            if (L.h_buf[r][k]) cudaFreeHost(L.h_buf[r][k]);
            L.d_buf[r][k] = nullptr; L.h_buf[r][k] = nullptr;
        }
    L.slabs = false;
    // device slabs now; pinned slabs lazily (only pageable operands need them)
    for (int r = 0; r < kRing; ++r)
        for (int k = 0; k < 3; ++k) PNCU_CUDA(cudaMalloc((void**)&L.d_buf[r][k], g.slab_bytes));
    L.slab_bytes = g.slab_bytes;
    L.slabs = true;
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

// a host scalar operand (NumPy's `a + 5.0`): one tiny copy on the compute stream; returns the device address
int stage_scalar(Lane& L, const char* host, int isz, int which, const char** dptr) {
    const int off = which == 0 ? kOffScalar1 : kOffScalar2;
    memcpy(L.h_small + off, host, isz);
    PNCU_CUDA(cudaMemcpyAsync(L.d_scalar + off, L.h_small + off, isz, cudaMemcpyHostToDevice, L.s_k));
    *dptr = L.d_scalar + off;
    return 0;
}

// ---- managed operands: where do the pages live? ------------------------------------------------------------
// A kernel that touches managed pages which are not resident on its GPU demand-faults them over PCIe at a
// fraction of the copy rate, so a managed range is prefetched to the GPU that is about to use it.  Prefetching
// a range that is already there still costs ~1 ms of driver time per GiB (page-table walk), 5x the kernel it
// precedes.  So a range is prefetched the FIRST time it is seen in a given placement and trusted afterwards
// (the note lives with the allocation in the registry); if host code touched some pages in between, the kernel
// faults those back (correct, slower, and then they are resident again).
std::mutex g_note_mu;

// `where`: GPU index for a whole range; 100 + G with `piece` for "sharded over G GPUs in pieces of `piece` bytes"
bool note_hit(const Operand& o, const char* lo, size_t bytes, int where, size_t piece) {
    if (!o.notes) return false;
    for (const ResidencyNote& n : *o.notes)
        if (n.where == where && n.piece == piece && n.lo <= lo && lo + bytes <= n.lo + n.bytes &&
            (where < 100 || (n.lo == lo && n.bytes == bytes)))
            return true;
    return false;
}
void note_set(const Operand& o, const char* lo, size_t bytes, int where, size_t piece) {
    if (!o.notes) return;
    std::vector<ResidencyNote>& v = *o.notes;
    for (size_t i = 0; i < v.size();) {   // whatever was recorded for overlapping bytes is stale now
        if (v[i].lo < lo + bytes && lo < v[i].lo + v[i].bytes) { v[i] = v.back(); v.pop_back(); }
        else ++i;
    }
    if (v.size() >= 32) v.clear();
    ResidencyNote n = {lo, bytes, where, piece};
    v.push_back(n);
}
// the GPU a managed range was last placed on as a whole (0 when unknown)
int home_of(const Operand& o, const char* lo, size_t bytes) {
    if (o.kind == MK_DEVICE && o.device >= 0 && o.device < g.ndev) return o.device;
    if (o.notes) {
        std::lock_guard<std::mutex> lk(g_note_mu);
        for (const ResidencyNote& n : *o.notes)
            if (n.where < 100 && n.where < g.ndev && n.lo <= lo && lo + bytes <= n.lo + n.bytes) return n.where;
    }
    return 0;
}
// The HEAD PAGE of every managed block we hand out stays in host memory and is mapped into every GPU
// (cudaMemAdviseSetPreferredLocation = CPU + SetAccessedBy each GPU, see nte_alloc_managed): NumPy reads element 0
// of an array ON THE HOST before every min / max reduction (it seeds the accumulator with a[0] and passes a[1:] to
// the loop).  If that page lived on the GPU, every a.max() would fault it to the host and the kernel would fault it
// back (~50 us per call, measured); as it is, the host reads it locally and the kernels read / write its 4 KiB
// over PCIe.  Prefetches therefore skip it.
constexpr size_t kHeadBytes = 4096;
int prefetch_clipped(const Operand& o, const char* lo, size_t bytes, Lane& L) {
    if (o.base && lo < o.base + kHeadBytes) {
        const size_t skip = (size_t)(o.base + kHeadBytes - lo);
        if (skip >= bytes) return 0;
        lo += skip;
        bytes -= skip;
    }
    PNCU_CUDA(cudaMemPrefetchAsync(lo, bytes, L.device, L.s_k));
    g_stats.prefetches += 1;
    return 0;
}
// make [lo, lo + bytes) of a managed operand resident on lane L's GPU (no-op for device memory / known placement)
int place_whole(const Operand& o, const char* lo, size_t bytes, Lane& L) {
    if (o.kind != MK_MANAGED || !bytes) return 0;
    std::lock_guard<std::mutex> lk(g_note_mu);
    if (note_hit(o, lo, bytes, L.device, 0)) return 0;
    PNCU_CUDA(cudaSetDevice(L.device));
    int rc = prefetch_clipped(o, lo, bytes, L);
    if (rc) return rc;
    note_set(o, lo, bytes, L.device, 0);
    return 0;
}

// ---- shard plans ----------------------------------------------------------------------------------------
// `lanes` contiguous shards at multiples of 65536 elements, as even as the unit allows (no empty shard)
void cut_shards(Job& j, int lanes, int64_t align = kAlign) {
    const int64_t kAlign = align;   // (shadows the default unit: reductions over several GPUs cut at whole fold groups)
    const int64_t units = (j.len + kAlign - 1) / kAlign;
    if (lanes > units) lanes = (int)units;
    if (lanes < 1) lanes = 1;
    j.nlanes = lanes;
    const int64_t base = units / lanes, extra = units % lanes;
    int64_t u = 0;
    for (int l = 0; l <= lanes; ++l) {
        const int64_t e = u * kAlign;
        j.shard[l] = e < j.len ? e : j.len;
        u += base + (l < extra ? 1 : 0);
    }
    j.shard[lanes] = j.len;
    for (int l = 0; l < lanes; ++l) { j.rc[l] = 0; j.err[l][0] = 0; j.lane_id[l] = l; }
}

// lanes for a call on HOST operands: min(op_max_workers, workers) + 1, never more than one per 65536-element unit
int plan(Job& j, int max_workers) {
    int lanes = 1;
    if (g.threading) {
        int extra = max_workers < g.workers ? max_workers : g.workers;
        if (extra < 0) extra = 0;
        lanes = extra + 1;
    }
    if (lanes > kMaxLanes) lanes = kMaxLanes;
    cut_shards(j, lanes);
    g_stats.lanes_last = j.nlanes;
    g_stats.gpus_last = j.nlanes < g.ndev ? j.nlanes : g.ndev;
    for (int l = 0; l < j.nlanes; ++l) {
        int rc = ensure_slabs(l);
        if (rc) return rc;
    }
    return 0;
}
// one lane, on the GPU that owns a device-memory operand
int plan_on_device(Job& j, int dev, bool slabs) {
    cut_shards(j, 1);
    j.lane_id[0] = dev;
    g_stats.lanes_last = 1;
    g_stats.gpus_last = 1;
    return slabs ? ensure_slabs(dev) : ensure_lane(dev);
}
// resident operands: one lane per GPU
int plan_one_lane_per_gpu(Job& j, int64_t align = kAlign) {
    cut_shards(j, g.ndev, align);
    for (int l = 0; l < j.nlanes; ++l) {
        int rc = ensure_lane(l);
        if (rc) return rc;
    }
    g_stats.lanes_last = j.nlanes;
    g_stats.gpus_last = j.nlanes;
    return 0;
}

int first_error(Job& j) {
    for (int l = 0; l < j.nlanes; ++l)
        if (j.rc[l]) { set_error("lane %d (GPU %d): %s", l, g.lanes[j.lane_id[l]].device, j.err[l]); return j.rc[l]; }
    return 0;
}

int decline() { g_stats.declined += 1; return NTE_DECLINED; }

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
int64_t lowest_threshold() { return threshold_for(NTE_CLASS_HEAVY, false, 8); }

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

// ---- one lane of an element-wise job with at least one HOST operand -----------------------------------------
// Host operands go slab by slab through the H2D -> kernel -> D2H pipeline; operands that are RESIDENT (managed or
// device memory next to host arrays) are used in place by the same kernels and cross no bus.
void lane_elementwise(Lane& L, Job& j, int li) {
    if (cudaSetDevice(L.device) != cudaSuccess) { cudaGetLastError(); fail(j, li, PNCU_ERR_NO_DEVICE); return; }
    const int64_t e0 = j.shard[li], e1 = j.shard[li + 1];
    if (e0 >= e1) return;
    const bool binary = j.kind == JOB_BINARY;
    const int64_t wide = j.isz_in > j.isz_out ? j.isz_in : j.isz_out;
    const bool vec1 = j.s1 != 0, vec2 = binary && j.s2 != 0;
    const bool res1 = j.k1.resident(), res2 = binary && j.k2.resident(), reso = j.ko.resident();
    // the same array twice (np.multiply(a, a); reference: in1 == in2 loads once, src/atop/ops_binary.cpp:631-658)
    const bool same12 = vec1 && vec2 && !res1 && j.in1 == j.in2 && j.s1 == j.s2;
    const bool out_strided = j.so != j.isz_out;
    const bool stage_out = !reso && (j.ko.kind == MK_PAGEABLE || out_strided);
    const bool stage_in = (vec1 && !res1 && (j.k1.kind == MK_PAGEABLE || j.s1 != j.isz_in)) ||
                          (vec2 && !res2 && !same12 && (j.k2.kind == MK_PAGEABLE || j.s2 != j.isz_in));
    const int64_t slab_elems = slab_elems_for(L.slab_bytes, wide, e1 - e0, stage_in || stage_out);
    const int64_t nslabs = (e1 - e0 + slab_elems - 1) / slab_elems;
    int rc;

    // broadcast scalars: one tiny copy per call (a resident scalar is read in place)
    const char* d_s1 = j.in1;
    const char* d_s2 = j.in2;
    if (!vec1 && !res1 && (rc = stage_scalar(L, j.in1, j.isz_in, 0, &d_s1))) { fail(j, li, rc); return; }
    if (binary && !vec2 && !res2 && (rc = stage_scalar(L, j.in2, j.isz_in, 1, &d_s2))) { fail(j, li, rc); return; }
    // resident array operands: this lane's shard onto this lane's GPU
    {
        const char* lo;
        int64_t bytes;
        if (vec1 && res1) {
            extent_of(j.in1 + e0 * j.s1, e1 - e0, j.s1, j.isz_in, &lo, &bytes);
            if ((rc = place_whole(j.k1, lo, (size_t)bytes, L))) { fail(j, li, rc); return; }
        }
        if (vec2 && res2) {
            extent_of(j.in2 + e0 * j.s2, e1 - e0, j.s2, j.isz_in, &lo, &bytes);
            if ((rc = place_whole(j.k2, lo, (size_t)bytes, L))) { fail(j, li, rc); return; }
        }
        if (reso) {
            extent_of(j.out + e0 * j.so, e1 - e0, j.so, j.isz_out, &lo, &bytes);
            if ((rc = place_whole(j.ko, lo, (size_t)bytes, L))) { fail(j, li, rc); return; }
        }
    }

    auto slab_range = [&](int64_t i, int64_t* off, int64_t* cnt) {
        *off = e0 + i * slab_elems;
        *cnt = (e1 - *off) < slab_elems ? (e1 - *off) : slab_elems;
    };
    auto retire = [&](int64_t i) -> int {  // slab i's result is where it belongs once its last event fires
        const int r = (int)(i % kRing);
        cudaError_t e = cudaEventSynchronize(reso ? L.ev_k[r] : L.ev_d2h[r]);
        if (e != cudaSuccess) return cuda_fail(e, "cudaEventSynchronize(slab)", __FILE__, __LINE__);
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
        if (i >= kRing && (rc = retire(i - kRing))) { fail(j, li, rc); return; }
        int64_t off, cnt;
        slab_range(i, &off, &cnt);
        const size_t in_bytes = (size_t)(cnt * j.isz_in), out_bytes = (size_t)(cnt * j.isz_out);
        // ---- host inputs -> device slab
        for (int k = 0; k < 2; ++k) {
            if (k == 0 ? (!vec1 || res1) : (!vec2 || res2 || same12)) continue;
            const int64_t stride = k == 0 ? j.s1 : j.s2;
            const char* src = (k == 0 ? j.in1 : j.in2) + off * stride;
            const int pk = k == 0 ? j.k1.kind : j.k2.kind;
            if (pk == MK_PAGEABLE || stride != j.isz_in) {
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
        const char* p1 = !vec1 ? d_s1 : (res1 ? j.in1 + off * j.s1 : L.d_buf[r][0]);
        const int64_t st1 = !vec1 ? 0 : (res1 ? j.s1 : j.isz_in);
        char* po = reso ? j.out + off * j.so : L.d_buf[r][2];
        const int64_t sto = reso ? j.so : j.isz_out;
        if (binary) {
            const char* p2 = !vec2 ? d_s2 : (res2 ? j.in2 + off * j.s2 : (same12 ? L.d_buf[r][0] : L.d_buf[r][1]));
            const int64_t st2 = !vec2 ? 0 : (res2 ? j.s2 : j.isz_in);
            rc = j.fn2(p1, p2, po, cnt, st1, st2, sto, (pncu_stream_t)L.s_k);
        } else {
            rc = j.fn1(p1, po, cnt, st1, sto, (pncu_stream_t)L.s_k);
        }
        if (rc) { fail(j, li, rc); return; }
        g_stats.launches += 1;
        LCUDA(cudaEventRecord(L.ev_k[r], L.s_k));
        // the next H2D into this slot must not overtake this kernel's reads
        LCUDA(cudaStreamWaitEvent(L.s_h2d, L.ev_k[r], 0));
        if (reso) continue;
        // ---- result -> host
        LCUDA(cudaStreamWaitEvent(L.s_d2h, L.ev_k[r], 0));
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
        if ((rc = retire(i))) { fail(j, li, rc); return; }
    }
}

// ---- one lane of a reduction / arg-reduction over a HOST array: slabs H2D -> pass-1 partials on this lane's GPU
void lane_reduce(Lane& L, Job& j, int li) {
    if (cudaSetDevice(L.device) != cudaSuccess) { cudaGetLastError(); fail(j, li, PNCU_ERR_NO_DEVICE); return; }
    const int64_t e0 = j.shard[li], e1 = j.shard[li + 1];
    if (e0 >= e1) return;
    const int64_t nslots = j.slot0[li + 1] - j.slot0[li];
    int rc;
    if ((rc = ensure_partials(L, (size_t)nslots * j.partial_bytes))) { fail(j, li, rc); return; }
    const bool staged = j.k1.kind == MK_PAGEABLE || j.s1 != j.isz_in;
    const int64_t slab_elems = slab_elems_for(L.slab_bytes, j.isz_in, e1 - e0, staged);
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
        if (staged) {
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

// ---- one lane of the fused chain over HOST arrays: three input slabs -> a*b+c (-> D2H) or -> block partials of its sum
void lane_fused(Lane& L, Job& j, int li) {
    if (cudaSetDevice(L.device) != cudaSuccess) { cudaGetLastError(); fail(j, li, PNCU_ERR_NO_DEVICE); return; }
    const int64_t e0 = j.shard[li], e1 = j.shard[li + 1];
    if (e0 >= e1) return;
    const bool to_sum = j.kind == JOB_MULADD_SUM;
    const int isz = j.isz_in;
    int rc;
    if ((rc = ensure_third_input(L))) { fail(j, li, rc); return; }
    if (to_sum && (rc = ensure_partials(L, (size_t)(j.slot0[li + 1] - j.slot0[li]) * j.partial_bytes))) { fail(j, li, rc); return; }
    const bool out_pageable = !to_sum && j.ko.kind == MK_PAGEABLE;
    const int64_t slab_elems = slab_elems_for(L.slab_bytes, isz, e1 - e0,
                                              j.k1.kind == MK_PAGEABLE || j.k2.kind == MK_PAGEABLE || j.k3.kind == MK_PAGEABLE || out_pageable);
    const int64_t nslabs = (e1 - e0 + slab_elems - 1) / slab_elems;
    const char* ins[3] = {j.in1, j.in2, j.in3};
    const int pks[3] = {j.k1.kind, j.k2.kind, j.k3.kind};
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
        if (out_pageable) {
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
        int dslot[3];
        for (int k = 0; k < 3; ++k) {
            // an operand passed twice (a*a + c) is uploaded once
            dslot[k] = slot_of[k];
            bool dup = false;
            for (int q = 0; q < k; ++q)
                if (ins[q] == ins[k]) { dslot[k] = dslot[q]; dup = true; break; }
            if (dup) continue;
            const char* src = ins[k] + off * isz;
            if (pks[k] == MK_PAGEABLE) {
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
            rc = pncu_muladd_sum_partials(j.atype, L.d_buf[r][dslot[0]], L.d_buf[r][dslot[1]], L.d_buf[r][dslot[2]], cnt,
                                          L.d_partials + slot * j.partial_bytes, (pncu_stream_t)L.s_k);
        } else {
            rc = j.fn3(L.d_buf[r][dslot[0]], L.d_buf[r][dslot[1]], L.d_buf[r][dslot[2]], L.d_buf[r][2], cnt, (pncu_stream_t)L.s_k);
        }
        if (rc) { fail(j, li, rc); return; }
        g_stats.launches += 1;
        LCUDA(cudaEventRecord(L.ev_k[r], L.s_k));
        LCUDA(cudaStreamWaitEvent(L.s_h2d, L.ev_k[r], 0));
        if (!to_sum) {
            LCUDA(cudaStreamWaitEvent(L.s_d2h, L.ev_k[r], 0));
            char* dst = j.out + off * isz;
            if (out_pageable) {
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
struct Post { Job* job; LaneFn fn; int logical; };
Post g_post[kMaxLanes];

void worker_thunk(void* ctx, int physical_lane) {
    Post* p = (Post*)ctx;
    p->fn(g.lanes[physical_lane], *p->job, p->logical);
}

// one wake, one wait (WorkMain, src/atop/threads.h:1079-1168): post every other lane to its worker, run lane 0
// on the calling thread, then wait for the workers
void run_lanes(Job& j, LaneFn fn) {
    for (int l = 1; l < j.nlanes; ++l) {
        const int p = j.lane_id[l];
        nte::Worker*& w = g.pool[p];
        if (!w) {
            w = new nte::Worker();
            w->start(p, g.spin_ns);
        }
        g_post[p].job = &j; g_post[p].fn = fn; g_post[p].logical = l;
        w->post(worker_thunk, &g_post[p]);
    }
    TRACE(TR_POSTED);
    fn(g.lanes[j.lane_id[0]], j, 0);  // the caller is lane 0
    TRACE(TR_LAUNCHED);
    for (int l = 1; l < j.nlanes; ++l) g.pool[j.lane_id[l]]->wait();
    TRACE(TR_JOINED);
}

// ---- RESIDENT element-wise work: one launch per GPU, in place -------------------------------------------------
// Operands in managed / device memory.  Shard l of a large managed array lives on GPU l (placed there the first
// time it is seen, then remembered); every call runs one kernel per GPU on its own shard and waits once.
constexpr size_t kShardResidentMinBytes = (size_t)64 << 20;  // smaller arrays stay on one GPU

void lane_resident(Lane& L, Job& j, int li) {
    if (cudaSetDevice(L.device) != cudaSuccess) { cudaGetLastError(); fail(j, li, PNCU_ERR_NO_DEVICE); return; }
    const int64_t e0 = j.shard[li], e1 = j.shard[li + 1];
    if (e0 >= e1) return;
    const int64_t n = e1 - e0;
    int rc;
    if (j.kind == JOB_MULADD) {
        rc = j.fn3(j.in1 + e0 * j.isz_in, j.in2 + e0 * j.isz_in, j.in3 + e0 * j.isz_in, j.out + e0 * j.isz_in, n, (pncu_stream_t)L.s_k);
    } else {
        const bool binary = j.kind == JOB_BINARY;
        const char* in1 = j.in1 + e0 * j.s1;
        const char* in2 = binary ? j.in2 + e0 * j.s2 : nullptr;
        char* out = j.out + e0 * j.so;
        if (j.s1 == 0 && !j.k1.resident() && (rc = stage_scalar(L, j.in1, j.isz_in, 0, &in1))) { fail(j, li, rc); return; }
        if (binary && j.s2 == 0 && !j.k2.resident() && (rc = stage_scalar(L, j.in2, j.isz_in, 1, &in2))) { fail(j, li, rc); return; }
        if (binary) rc = j.fn2(in1, in2, out, n, j.s1, j.s2, j.so, (pncu_stream_t)L.s_k);
        else rc = j.fn1(in1, out, n, j.s1, j.so, (pncu_stream_t)L.s_k);
    }
    if (rc) { fail(j, li, rc); return; }
    g_stats.launches += 1;
    cudaError_t e = cudaStreamSynchronize(L.s_k);
    if (e != cudaSuccess) { cuda_fail(e, "cudaStreamSynchronize(resident shard)", __FILE__, __LINE__); fail(j, li, (int)e); }
}

// place every managed array operand of a resident job: sharded over the job's lanes, or whole on its one lane.
// Runs on the calling thread before the lanes are posted (a hit costs one short list walk per operand).
int place_operands(Job& j) {
    struct Arr { const Operand* o; const char* p; int64_t stride; int isz; };
    Arr arrs[4];
    int na = 0;
    if (j.s1 != 0 && j.in1) arrs[na++] = Arr{&j.k1, j.in1, j.s1, j.isz_in};
    if ((j.kind == JOB_BINARY || j.kind == JOB_MULADD || j.kind == JOB_MULADD_SUM) && j.s2 != 0 && j.in2) arrs[na++] = Arr{&j.k2, j.in2, j.s2, j.isz_in};
    if ((j.kind == JOB_MULADD || j.kind == JOB_MULADD_SUM) && j.in3) arrs[na++] = Arr{&j.k3, j.in3, j.isz_in, j.isz_in};
    if (j.out && j.so != 0) arrs[na++] = Arr{&j.ko, j.out, j.so, j.isz_out};
    for (int a = 0; a < na; ++a) {
        const Arr& A = arrs[a];
        if (A.o->kind != MK_MANAGED) continue;
        if (j.nlanes == 1) {
            const char* lo;
            int64_t bytes;
            extent_of(A.p, j.len, A.stride, A.isz, &lo, &bytes);
            int rc = place_whole(*A.o, lo, (size_t)bytes, g.lanes[j.lane_id[0]]);
            if (rc) return rc;
            continue;
        }
        // sharded (contiguous operands only): note = (range, 100 + lanes, bytes of the first shard)
        const char* lo = A.p;
        const size_t bytes = (size_t)j.len * A.isz;
        const size_t piece = (size_t)(j.shard[1] - j.shard[0]) * A.isz;
        std::lock_guard<std::mutex> lk(g_note_mu);
        if (note_hit(*A.o, lo, bytes, 100 + j.nlanes, piece)) continue;
        for (int l = 0; l < j.nlanes; ++l) {
            Lane& L = g.lanes[j.lane_id[l]];
            const size_t b = (size_t)(j.shard[l + 1] - j.shard[l]) * A.isz;
            if (!b) continue;
            PNCU_CUDA(cudaSetDevice(L.device));
            int rc = prefetch_clipped(*A.o, A.p + j.shard[l] * A.isz, b, L);
            if (rc) return rc;
        }
        note_set(*A.o, lo, bytes, 100 + j.nlanes, piece);
    }
    return 0;
}

// are all array operands contiguous managed memory, and is the job big enough to spread over the GPUs?
bool want_sharding(const Job& j, int narrays, const Operand* const* ops, const int64_t* strides, const int* iszs) {
    if (g.ndev < 2 || !g.threading) return false;
    for (int a = 0; a < narrays; ++a)
        if (ops[a]->kind != MK_MANAGED || strides[a] != iszs[a]) return false;
    const int wide = j.isz_in > j.isz_out ? j.isz_in : j.isz_out;
    return (size_t)j.len * (size_t)wide >= kShardResidentMinBytes;
}

// ---- RESIDENT reductions: one kernel per GPU, exchange and fold on the device --------------------------------
// Each lane launches ONE kernel (pncu_reduce_fused / pncu_argminmax_fused / pncu_muladd_sum_fused): pass 1 on its
// shard; the last CTA of a non-root GPU stores the shard's block partials into GPU 0's table over NVLink and
// bumps its arrival counter; GPU 0's last CTA waits for the other shards, folds the whole table in slot order,
// writes the result into lane 0's MAPPED host scratch and then the completion flag, which the caller polls.
int launch_fused_shard(Lane& L, Job& j, int li, const pncu_fused_exchange* x, void* out, const void* start) {
    const int64_t e0 = j.shard[li], n = j.shard[li + 1] - j.shard[li];
    int rc;
    if (j.kind == JOB_REDUCE)
        rc = pncu_reduce_fused(j.func, j.atype, j.in1 + e0 * j.s1, n, x, out, start, (pncu_stream_t)L.s_k);
    else if (j.kind == JOB_ARG)
        rc = pncu_argminmax_fused(j.argkind, j.atype, j.in1 + e0 * j.s1, n, e0, x, (int64_t*)out, (pncu_stream_t)L.s_k);
    else
        rc = pncu_muladd_sum_fused(j.atype, j.in1 + e0 * j.isz_in, j.in2 + e0 * j.isz_in, j.in3 + e0 * j.isz_in, n, x, out,
                                   (pncu_stream_t)L.s_k);
    if (!rc) g_stats.launches += 1;
    return rc;
}

pncu_fused_exchange g_x[kMaxLanes];  // per-lane exchange descriptors of the call in flight

void lane_reduce_shard(Lane& L, Job& j, int li) {   // lanes >= 1: launch and return
    if (cudaSetDevice(L.device) != cudaSuccess) { cudaGetLastError(); fail(j, li, PNCU_ERR_NO_DEVICE); return; }
    int rc = launch_fused_shard(L, j, li, &g_x[li], nullptr, nullptr);
    if (rc) fail(j, li, rc);
}

// poll lane L's completion flag for `want`; the stream is queried now and then so that a failed launch or a
// faulted kernel surfaces as an error instead of an endless wait
int wait_flag(Lane& L, uint64_t want) {
    volatile unsigned long long* f = reinterpret_cast<volatile unsigned long long*>(L.h_small + kOffFlag);
    unsigned polls = 0;
    for (;;) {
        const unsigned long long v = __atomic_load_n(f, __ATOMIC_ACQUIRE);
        if ((v & ~PNCU_FUSED_TIMEOUT_BIT) == want) {
            if (v & PNCU_FUSED_TIMEOUT_BIT) {
                set_error("a GPU never delivered its block partials to the reduction exchange (waited ~2 s); no result was produced");
                return PNCU_ERR_TIMEOUT;
            }
            return 0;
        }
        NTE_PAUSE();
        if ((++polls & 0x3ff) == 0) {
            cudaError_t e = cudaStreamQuery(L.s_k);
            if (e == cudaErrorNotReady) continue;
            if (e != cudaSuccess) return cuda_fail(e, "cudaStreamQuery(reduction)", __FILE__, __LINE__);
            const unsigned long long v2 = __atomic_load_n(f, __ATOMIC_ACQUIRE);
            if ((v2 & ~PNCU_FUSED_TIMEOUT_BIT) == want) continue;   // published between the two looks
            set_error("the reduction kernel finished without publishing a result");
            return PNCU_ERR_ARGUMENT;
        }
    }
}

// bring the exchange back to a known state after a failure (all lanes idle first)
void reset_exchange(Job& j) {
    for (int l = 0; l < j.nlanes; ++l) {
        Lane& L = g.lanes[j.lane_id[l]];
        if (cudaSetDevice(L.device) != cudaSuccess) { cudaGetLastError(); continue; }
        if (cudaStreamSynchronize(L.s_k) != cudaSuccess) cudaGetLastError();
        pncu_fused_reset((pncu_stream_t)L.s_k);
        if (cudaStreamSynchronize(L.s_k) != cudaSuccess) cudaGetLastError();
    }
    Lane& L0 = g.lanes[0];
    if (L0.ready && cudaSetDevice(L0.device) == cudaSuccess) {
        if (cudaMemset(L0.d_scalar + kOffArrived, 0, 8) != cudaSuccess) cudaGetLastError();
    }
    g.x_expected = 0;
}

int run_reduce_resident(Job& j, bool sharded) {
    int rc;
    if (sharded) {
        // shards start at whole fold groups (64 slots = 4 MiB of input): every GPU folds its own groups and only the
        // group partials cross NVLink (pncu_fused_exchange::groups); any split gives the single-GPU bits
        const int64_t group_elems = (int64_t)pncu_reduce_group_slots() * (kBlockBytes / j.isz_in);
        if ((rc = plan_one_lane_per_gpu(j, group_elems > kAlign ? group_elems : kAlign))) return rc;
    } else {
        const char* lo;
        int64_t bytes;
        extent_of(j.in1, j.len, j.s1, j.isz_in, &lo, &bytes);
        if ((rc = plan_on_device(j, home_of(j.k1, lo, (size_t)bytes), false))) return rc;
    }
    TRACE(TR_PLANNED);
    if ((rc = place_operands(j))) return rc;
    TRACE(TR_PLACED);
    Lane& L0 = g.lanes[j.lane_id[0]];
    PNCU_CUDA(cudaSetDevice(L0.device));
    const void* start = nullptr;
    if (j.host_start) {
        memcpy(L0.h_small + kOffStart, j.host_start, j.out_bytes);
        start = L0.h_small + kOffStart;
    }
    const uint64_t want = ++L0.seq;
    if (j.nlanes == 1) {
        pncu_fused_exchange& x = g_x[0];
        memset(&x, 0, sizeof(x));
        x.n = 1;
        x.done_flag = reinterpret_cast<unsigned long long*>(L0.h_small + kOffFlag);
        x.done_value = want;
        if ((rc = launch_fused_shard(L0, j, 0, &x, L0.h_small + kOffResult, start))) return rc;
        if ((rc = wait_flag(L0, want))) { reset_exchange(j); return rc; }
    } else {
        for (int l = 0; l <= j.nlanes; ++l) j.slot0[l] = pncu_reduce_partial_count(j.atype, j.shard[l]);
        const size_t slots_bytes = (((size_t)j.slot0[j.nlanes] * j.partial_bytes + 255) / 256) * 256;
        const int group = pncu_reduce_group_slots();
        const size_t need = slots_bytes + (size_t)((j.slot0[j.nlanes] + group - 1) / group) * j.partial_bytes;
        bool aligned = true;
        for (int l = 0; l < j.nlanes; ++l) aligned = aligned && j.slot0[l] % group == 0;
        if (g.x_table_bytes < need) {
            if (g.x_table) cudaFree(g.x_table);
            g.x_table = nullptr;
            g.x_table_bytes = 0;
            const size_t cap = need < ((size_t)1 << 20) ? ((size_t)1 << 20) : need * 2;
            PNCU_CUDA(cudaMalloc((void**)&g.x_table, cap));
            g.x_table_bytes = cap;
        }
        g.x_expected += (unsigned long long)(j.nlanes - 1);
        for (int l = 0; l < j.nlanes; ++l) {
            pncu_fused_exchange& x = g_x[l];
            memset(&x, 0, sizeof(x));
            x.n = j.nlanes; x.self = l; x.all = 0; x.root = 0;
            x.slot_base = j.slot0[l];
            x.total_count = j.slot0[j.nlanes];
            x.slots[0] = g.x_table;
            x.group_offset = aligned ? (int64_t)slots_bytes : 0;
            x.arrived[0] = reinterpret_cast<unsigned long long*>(L0.d_scalar + kOffArrived);
            x.expected = g.x_expected;
            if (l == 0) {
                x.done_flag = reinterpret_cast<unsigned long long*>(L0.h_small + kOffFlag);
                x.done_value = want;
            }
        }
        // post the other GPUs' launches, launch the root here, wait for the posts (launch status), then for the flag
        for (int l = 1; l < j.nlanes; ++l) {
            const int p = j.lane_id[l];
            nte::Worker*& w = g.pool[p];
            if (!w) { w = new nte::Worker(); w->start(p, g.spin_ns); }
            g_post[p].job = &j; g_post[p].fn = lane_reduce_shard; g_post[p].logical = l;
            w->post(worker_thunk, &g_post[p]);
        }
        TRACE(TR_POSTED);
        rc = launch_fused_shard(L0, j, 0, &g_x[0], L0.h_small + kOffResult, start);
        if (rc) fail(j, 0, rc);
        TRACE(TR_LAUNCHED);
        for (int l = 1; l < j.nlanes; ++l) g.pool[j.lane_id[l]]->wait();
        TRACE(TR_JOINED);
        if ((rc = first_error(j))) { reset_exchange(j); return rc; }   // a shard was never launched: do not wait for it
        if ((rc = wait_flag(L0, want))) { reset_exchange(j); return rc; }
    }
    memcpy(j.host_out, L0.h_small + kOffResult, j.out_bytes);
    g_stats.d2h += j.out_bytes;
    TRACE(TR_DONE);
    return 0;
}

// ---- host-operand reductions: gather every lane's partials on GPU 0, fold there in fixed order ----------------
int finish_on_gpu0(Job& j) {
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
            Lane& L = g.lanes[j.lane_id[l]];
            // peer copy (NVLink P2P when enabled, else staged by the driver); same-GPU lanes: plain D2D
            PNCU_CUDA(cudaMemcpyPeerAsync(all + j.slot0[l] * j.partial_bytes, L0.device, L.d_partials, L.device,
                                          (size_t)cnt * j.partial_bytes, L0.s_k));
        }
    }
    // the finish kernel reads the start value from, and writes the result into, lane 0's mapped host scratch
    const void* start = nullptr;
    if (j.host_start) {
        memcpy(L0.h_small + kOffStart, j.host_start, j.out_bytes);
        start = L0.h_small + kOffStart;
    }
    int rc;
    if (j.kind == JOB_ARG) rc = pncu_argminmax_finish(j.argkind, j.atype, all, total, (int64_t*)(L0.h_small + kOffResult), (pncu_stream_t)L0.s_k);
    else rc = pncu_reduce_finish(j.func, j.atype, all, total, L0.h_small + kOffResult, start, (pncu_stream_t)L0.s_k);
    if (rc) return rc;
    g_stats.launches += 1;
    PNCU_CUDA(cudaStreamSynchronize(L0.s_k));
    memcpy(j.host_out, L0.h_small + kOffResult, j.out_bytes);
    g_stats.d2h += j.out_bytes;
    return 0;
}

bool usable() { return g.inited && !g.forked.load(std::memory_order_relaxed); }

// fault injection for the error-path tests: the n-th GPU dispatch from now reports a launch failure
int injected_fault() {
    if (g.fault_countdown > 0 && --g.fault_countdown == 0) {
        set_error("injected fault (nte_inject_fault / PNUMPY_CUDA_FAULT_INJECT): simulated cudaErrorLaunchFailure");
        return (int)cudaErrorLaunchFailure;
    }
    return 0;
}

void atfork_child() { g.forked.store(true); }

}  // namespace

// =======================================================================================================
extern "C" {

int nte_init(void) {
    std::lock_guard<std::mutex> lk(g.api_mu);
    if (g.forked) { set_error("CUDA cannot be used in a fork()ed child; start workers with the 'spawn' method"); return PNCU_ERR_NO_DEVICE; }
    if (g.inited) return 0;
    int n = 0;
    int rc = pncu_init(&n);
    if (rc) return rc;
    DeviceGuard guard;
    g.ndev = n > kMaxLanes ? kMaxLanes : n;
    if (g.ndev > PNCU_MAX_PEERS) g.ndev = PNCU_MAX_PEERS;
    const char* env = getenv("PNUMPY_CUDA_MIN_ELEMS");
    if (env && *env) g.min_elems = atoll(env);
    env = getenv("PNUMPY_CUDA_SLAB_BYTES");
    if (env && *env) g.slab_bytes = (size_t)atoll(env);
    env = getenv("PNUMPY_CUDA_DEVICES");  // use only the first k devices
    if (env && *env && atoi(env) > 0 && atoi(env) < g.ndev) g.ndev = atoi(env);
    env = getenv("PNUMPY_CUDA_SPIN_US");
    if (env && *env) g.spin_ns = (uint64_t)atoll(env) * 1000ull;
    env = getenv("PNUMPY_CUDA_FAULT_INJECT");
    if (env && *env) g.fault_countdown = atoll(env);
    // peer access: the reduction exchange stores into GPU 0's table, the host path gathers partials with peer copies
    g.peer_all = true;
    for (int a = 0; a < g.ndev; ++a)
        for (int b = 0; b < g.ndev; ++b) {
            if (a == b) continue;
            int can = 0;
            bool ok = false;
            if (cudaDeviceCanAccessPeer(&can, a, b) == cudaSuccess && can) {
                cudaSetDevice(a);
                cudaError_t e = cudaDeviceEnablePeerAccess(b, 0);
                ok = e == cudaSuccess || e == cudaErrorPeerAccessAlreadyEnabled;
                if (e != cudaSuccess) cudaGetLastError();
            } else {
                cudaGetLastError();
            }
            if (!ok && b == 0) g.peer_all = false;
        }
    if (pncu_reduce_partial_count(PNCU_DOUBLE, kBlockBytes / 8) != 1 || pncu_reduce_partial_count(PNCU_DOUBLE, kBlockBytes / 8 + 1) != 2 ||
        pncu_reduce_block_elems() != kAlign) {
        set_error("nte: reduction block geometry differs from libpnumpy_cuda's");
        return PNCU_ERR_ARGUMENT;
    }
    static bool atfork_registered = false;
    if (!atfork_registered) { pthread_atfork(nullptr, nullptr, atfork_child); atfork_registered = true; }
    g.inited = true;
    return 0;
}

int nte_shutdown(void) {
    std::lock_guard<std::mutex> lk(g.api_mu);
    if (g.forked) return 0;   // the child owns neither the threads nor the CUDA context
    DeviceGuard guard;
    for (int l = 1; l < kMaxLanes; ++l) {
        nte::Worker* w = g.pool[l];
        if (!w) continue;
        w->stop();
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
    if (g.x_table) { cudaSetDevice(0); cudaFree(g.x_table); g.x_table = nullptr; g.x_table_bytes = 0; }
    g.x_expected = 0;
    g.inited = false;
    return 0;
}

int nte_device_count(void) { return g.inited ? g.ndev : 0; }
int nte_is_forked_child(void) { return g.forked ? 1 : 0; }

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
void nte_set_trace(int on) { g_trace_on = on != 0; }
void nte_get_trace(int64_t* out8) {
    if (!out8) return;
    for (int i = 0; i < TR_N; ++i) out8[i] = g_trace[i];
}
int64_t nte_inject_fault(int64_t nth) {
    std::lock_guard<std::mutex> lk(g.api_mu);
    int64_t p = g.fault_countdown;
    g.fault_countdown = nth < 0 ? 0 : nth;
    return p;
}

void* nte_alloc_pinned(size_t bytes) {
    if (g.forked || pncu_init(nullptr)) return nullptr;
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    registry_add(p, bytes ? bytes : 1, MK_PINNED, -1);
    return p;
}
void nte_free_pinned(void* ptr) {
    if (!ptr) return;
    registry_remove(ptr);
    if (g.forked) return;   // not ours to free in a fork()ed child
    if (cudaFreeHost(ptr) != cudaSuccess) cudaGetLastError();
}

void* nte_alloc_managed(size_t bytes) {
    if (g.forked || pncu_init(nullptr)) return nullptr;
    void* p = nullptr;
    if (cudaMallocManaged(&p, bytes ? bytes : 1, cudaMemAttachGlobal) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    // head page: host-resident, mapped into every GPU (see kHeadBytes); failures only cost speed
    if (bytes >= 2 * kHeadBytes) {
        bool ok = cudaMemAdvise(p, kHeadBytes, cudaMemAdviseSetPreferredLocation, cudaCpuDeviceId) == cudaSuccess;
        for (int d = 0; ok && d < (g.ndev > 0 ? g.ndev : 1); ++d)
            ok = cudaMemAdvise(p, kHeadBytes, cudaMemAdviseSetAccessedBy, d) == cudaSuccess;
        if (!ok) cudaGetLastError();
    }
    registry_add(p, bytes ? bytes : 1, MK_MANAGED, -1);
    return p;
}
void nte_free_managed(void* ptr) {
    if (!ptr) return;
    {
        // a dispatch in flight may hold the block's residency notes: wait for it
        std::lock_guard<std::mutex> lk(g.api_mu);
        std::lock_guard<std::mutex> nk(g_note_mu);
        registry_remove(ptr);
    }
    if (g.forked) return;
    if (cudaFree(ptr) != cudaSuccess) cudaGetLastError();
}
// zero a managed block ON the GPU (np.zeros under the managed recycler must not drag the pages to the host)
int nte_zero_managed(void* ptr, size_t bytes) {
    if (!ptr || !bytes) return 0;
    if (!usable()) return PNCU_ERR_NO_DEVICE;
    DeviceGuard guard;
    PNCU_CUDA(cudaSetDevice(0));
    PNCU_CUDA(cudaMemsetAsync(ptr, 0, bytes, cudaStreamPerThread));
    PNCU_CUDA(cudaStreamSynchronize(cudaStreamPerThread));
    Operand o;
    if (lookup_registry(ptr, &o) && o.notes) {
        std::lock_guard<std::mutex> nk(g_note_mu);
        note_set(o, (const char*)ptr, bytes, 0, 0);
    }
    return 0;
}
// a recycled managed block is about to be handed to a new owner: what we knew about its pages stays valid
// (they are where the last kernel left them), so nothing to do; exported for symmetry / tests
void nte_forget_residency(void* ptr) {
    AllocInfo ai;
    if (!ptr || !registry_find(ptr, &ai) || !ai.notes) return;
    std::lock_guard<std::mutex> lk(g.api_mu);
    std::lock_guard<std::mutex> nk(g_note_mu);
    ai.notes->clear();
}

void nte_get_stats(nte_stats* o) {
    if (!o) return;
    o->calls = g_stats.calls; o->declined = g_stats.declined; o->launches = g_stats.launches;
    o->h2d_bytes = g_stats.h2d; o->d2h_bytes = g_stats.d2h; o->staged_bytes = g_stats.staged;
    o->lanes_last = g_stats.lanes_last; o->gpus_last = g_stats.gpus_last; o->prefetches = g_stats.prefetches;
}
void nte_reset_stats(void) {
    g_stats.calls = 0; g_stats.declined = 0; g_stats.launches = 0; g_stats.h2d = 0; g_stats.d2h = 0;
    g_stats.staged = 0; g_stats.lanes_last = 0; g_stats.gpus_last = 0; g_stats.prefetches = 0;
}

// `out` may alias an input exactly (every element is read before its result is written, by the same thread or
// through staging); any other overlap needs the sequential semantics of the caller's loop (NumPy's accumulate
// calls a binary loop as (out, in + s, out + s); reference: src/pnumpy/_pnumpy.cpp:743-745)
static bool partial_overlap(const Job& j, bool binary) {
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
        if (overlap && !(ins[k] == j.out && strides[k] == j.so && j.isz_in == j.isz_out)) return true;
    }
    return false;
}

static int dispatch_elementwise(Job& j, int max_workers) {
    if (!usable()) {
        if (g.forked) return decline();   // a fork()ed child runs NumPy's loops
        set_error("nte_init() has not succeeded");
        return PNCU_ERR_NO_DEVICE;
    }
    if (j.len <= 0) return decline();
    TRACE(TR_ENTER);
    const bool binary = j.kind == JOB_BINARY;
    // 1. memory kinds from the registry (no driver call); small calls on memory we did not hand out are declined
    //    right here -- that is every small NumPy call on ordinary arrays, and it costs three map lookups
    const bool r1 = lookup_registry(j.in1, &j.k1);
    const bool r2 = binary ? lookup_registry(j.in2, &j.k2) : true;
    const bool ro = lookup_registry(j.out, &j.ko);
    const bool known_resident = j.k1.resident() || (binary && j.k2.resident()) || j.ko.resident();
    if (!known_resident && j.len < lowest_threshold()) return decline();
    std::unique_lock<std::mutex> lk(g.api_mu, std::try_to_lock);
    if (!lk.owns_lock()) return decline();  // re-entrant / concurrent caller: run the old loop
    TRACE(TR_LOCKED);
    if (!r1) classify_driver(j.in1, &j.k1);
    if (binary && !r2) classify_driver(j.in2, &j.k2);
    if (!ro) classify_driver(j.out, &j.ko);
    if (!binary) j.k2 = Operand();
    const bool vec1 = j.s1 != 0, vec2 = binary && j.s2 != 0;
    if (!vec1 && !vec2) return decline();   // scalar op scalar
    if (!stride_ok(j.in1, j.s1, j.isz_in) || (binary && !stride_ok(j.in2, j.s2, j.isz_in)) || !stride_ok(j.out, j.so, j.isz_out))
        return decline();
    if (partial_overlap(j, binary)) return decline();
    const int narr = 1 + (vec1 ? 1 : 0) + (vec2 ? 1 : 0);
    const int nres = (j.ko.resident() ? 1 : 0) + (vec1 && j.k1.resident() ? 1 : 0) + (vec2 && j.k2.resident() ? 1 : 0);
    DeviceGuard guard;
    int rc;
    if (nres == narr) {
        // ---- every array operand is resident: one launch per GPU, in place, whatever the size
        const Operand* ops[3]; int64_t strides[3]; int iszs[3]; int na = 0;
        ops[na] = &j.ko; strides[na] = j.so; iszs[na++] = j.isz_out;
        if (vec1) { ops[na] = &j.k1; strides[na] = j.s1; iszs[na++] = j.isz_in; }
        if (vec2) { ops[na] = &j.k2; strides[na] = j.s2; iszs[na++] = j.isz_in; }
        if ((rc = injected_fault())) return rc;
        if (want_sharding(j, na, ops, strides, iszs)) {
            if ((rc = plan_one_lane_per_gpu(j))) return rc;
        } else {
            const char* lo;
            int64_t bytes;
            extent_of(j.out, j.len, j.so, j.isz_out, &lo, &bytes);
            int dev = home_of(j.ko, lo, (size_t)bytes);
            if (vec1 && j.k1.kind == MK_DEVICE) dev = home_of(j.k1, j.in1, 1);
            if (vec2 && j.k2.kind == MK_DEVICE) dev = home_of(j.k2, j.in2, 1);
            if (j.ko.kind == MK_DEVICE) dev = home_of(j.ko, j.out, 1);
            if ((rc = plan_on_device(j, dev, false))) return rc;
        }
        TRACE(TR_PLANNED);
        if ((rc = place_operands(j))) return rc;
        TRACE(TR_PLACED);
        run_lanes(j, lane_resident);
        rc = first_error(j);
        if (!rc) g_stats.calls += 1;
        TRACE(TR_DONE);
        return rc;
    }
    {
        // ---- at least one HOST array operand: worth a PCIe round trip?  (Also when another operand is resident: NumPy
        // calls the loops of np.sin(int16_array) with 8192-element cast buffers as input and the managed result array as
        // output -- 120 staged launches per million elements would be 4x slower than its own loop.)
        // an op whose output does not depend on its input (isnan / isinf / isfinite of integers): shipping the input
        // over PCIe to memset the output would be absurd -- NumPy's own loop is a memset already
        if (nres == 0 && j.op_class == NTE_CLASS_CONST) return decline();
        const bool pageable = j.ko.kind == MK_PAGEABLE || (vec1 && j.k1.kind == MK_PAGEABLE) || (vec2 && j.k2.kind == MK_PAGEABLE);
        if (j.len < threshold_for(j.op_class, pageable, j.isz_in > j.isz_out ? j.isz_in : j.isz_out)) return decline();
    }
    // ---- at least one host array (possibly next to resident ones): the staging pipeline
    if ((rc = injected_fault())) return rc;
    int dev = -1;
    if (j.ko.kind == MK_DEVICE) dev = home_of(j.ko, j.out, 1);
    else if (vec1 && j.k1.kind == MK_DEVICE) dev = home_of(j.k1, j.in1, 1);
    else if (vec2 && j.k2.kind == MK_DEVICE) dev = home_of(j.k2, j.in2, 1);
    if (dev >= 0) rc = plan_on_device(j, dev, true);   // device memory belongs to one GPU
    else rc = plan(j, max_workers);
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

static int dispatch_reduce(Job& j, int max_workers) {
    if (!usable()) {
        if (g.forked) return decline();
        set_error("nte_init() has not succeeded");
        return PNCU_ERR_NO_DEVICE;
    }
    if (j.len <= 0) return decline();
    TRACE(TR_ENTER);
    const bool r1 = lookup_registry(j.in1, &j.k1);
    if (!j.k1.resident() && j.len < lowest_threshold()) return decline();
    std::unique_lock<std::mutex> lk(g.api_mu, std::try_to_lock);
    if (!lk.owns_lock()) return decline();
    TRACE(TR_LOCKED);
    if (!r1) classify_driver(j.in1, &j.k1);
    DeviceGuard guard;
    int rc;
    if (j.k1.resident()) {
        // resident input: one kernel per GPU; a large contiguous managed array is sharded over all of them
        if ((rc = injected_fault())) return rc;
        const bool sharded = g.ndev > 1 && g.threading && g.peer_all && j.k1.kind == MK_MANAGED && j.s1 == j.isz_in &&
                             (size_t)j.len * j.isz_in >= kShardResidentMinBytes;
        rc = run_reduce_resident(j, sharded);
        if (!rc) g_stats.calls += 1;
        return rc;
    }
    if (j.len < threshold_for(NTE_CLASS_REDUCE, j.k1.kind == MK_PAGEABLE, j.isz_in)) return decline();
    if ((rc = injected_fault())) return rc;
    if ((rc = plan(j, max_workers))) return rc;
    for (int l = 0; l <= j.nlanes; ++l) j.slot0[l] = pncu_reduce_partial_count(j.atype, j.shard[l]);
    run_lanes(j, lane_reduce);
    if ((rc = first_error(j))) return rc;
    rc = finish_on_gpu0(j);
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
    j.host_out = out; j.out_bytes = isz; j.host_start = use_out_as_start ? out : nullptr;
    return dispatch_reduce(j, max_workers);
}

int nte_argminmax(int kind, int atype, const void* in, int64_t len, int64_t* out_index, int max_workers) {
    const int isz = pncu_itemsize(atype);
    if (!isz || !pncu_GetArgMinMaxOpFast(kind, atype)) return decline();
    Job j;
    j.kind = JOB_ARG; j.argkind = kind; j.atype = atype; j.isz_in = isz; j.isz_out = 8; j.partial_bytes = 16;
    j.in1 = (const char*)in; j.len = len; j.s1 = isz;
    j.host_out = out_index; j.out_bytes = 8;
    return dispatch_reduce(j, max_workers);
}

// fused chain: element-wise (out != NULL) or summed (host_out)
static int dispatch_fused(Job& j, int max_workers) {
    if (!usable()) {
        if (g.forked) return decline();
        set_error("nte_init() has not succeeded");
        return PNCU_ERR_NO_DEVICE;
    }
    if (j.len <= 0) return decline();
    const bool to_sum = j.kind == JOB_MULADD_SUM;
    const bool r1 = lookup_registry(j.in1, &j.k1), r2 = lookup_registry(j.in2, &j.k2), r3 = lookup_registry(j.in3, &j.k3);
    const bool ro = to_sum ? true : lookup_registry(j.out, &j.ko);
    const bool known_resident = j.k1.resident() || j.k2.resident() || j.k3.resident() || (!to_sum && j.ko.resident());
    if (!known_resident && j.len < lowest_threshold()) return decline();
    std::unique_lock<std::mutex> lk(g.api_mu, std::try_to_lock);
    if (!lk.owns_lock()) return decline();
    if (!r1) classify_driver(j.in1, &j.k1);
    if (!r2) classify_driver(j.in2, &j.k2);
    if (!r3) classify_driver(j.in3, &j.k3);
    if (!ro) classify_driver(j.out, &j.ko);
    if (to_sum) j.ko = j.k1;
    const Operand* ks[4] = {&j.k1, &j.k2, &j.k3, &j.ko};
    int nhost = 0, npageable = 0;
    for (int k = 0; k < 4; ++k) { nhost += ks[k]->host(); npageable += ks[k]->kind == MK_PAGEABLE; }
    if (nhost != 0 && nhost != 4) return decline();  // host and resident arrays mixed: the caller's ufunc chain
    const size_t bytes = (size_t)j.len * j.isz_in;
    if (!to_sum) {
        // `out` may alias an input exactly; partial overlap is declined
        const char* ins[3] = {j.in1, j.in2, j.in3};
        for (int k = 0; k < 3; ++k) {
            const bool overlap = ins[k] < j.out + bytes && j.out < ins[k] + bytes;
            if (overlap && ins[k] != j.out) return decline();
        }
    }
    DeviceGuard guard;
    int rc;
    if (nhost == 0) {
        if ((rc = injected_fault())) return rc;
        const bool all_managed = j.k1.kind == MK_MANAGED && j.k2.kind == MK_MANAGED && j.k3.kind == MK_MANAGED && j.ko.kind == MK_MANAGED;
        const bool sharded = g.ndev > 1 && g.threading && all_managed && bytes >= kShardResidentMinBytes && (!to_sum || g.peer_all);
        if (to_sum) {
            rc = run_reduce_resident(j, sharded);
        } else {
            if (sharded) rc = plan_one_lane_per_gpu(j);
            else rc = plan_on_device(j, home_of(j.ko, j.out, bytes), false);
            if (rc) return rc;
            if ((rc = place_operands(j))) return rc;
            run_lanes(j, lane_resident);
            rc = first_error(j);
        }
        if (!rc) g_stats.calls += 1;
        return rc;
    }
    if (j.len < threshold_for(NTE_CLASS_STREAM, npageable != 0, j.isz_in)) return decline();
    if ((rc = injected_fault())) return rc;
    if ((rc = plan(j, max_workers))) return rc;
    if (to_sum)
        for (int l = 0; l <= j.nlanes; ++l) j.slot0[l] = pncu_reduce_partial_count(j.atype, j.shard[l]);
    run_lanes(j, lane_fused);
    if ((rc = first_error(j))) return rc;
    if (to_sum) {
        j.kind = JOB_REDUCE;  // the partials are those of the ADD reduction: ordinary gather + finish
        j.func = PNCU_ADD;
        if ((rc = finish_on_gpu0(j))) return rc;
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
    j.s1 = j.s2 = j.so = j.isz_in;
    return dispatch_fused(j, max_workers);
}

int nte_muladd_sum(int atype, const void* in1, const void* in2, const void* in3, void* out, int64_t len, int max_workers) {
    if (!pncu_GetMulAddSumFast(atype)) return decline();
    Job j;
    j.kind = JOB_MULADD_SUM; j.atype = atype; j.isz_in = j.isz_out = pncu_itemsize(atype);
    j.partial_bytes = pncu_reduce_partial_bytes(PNCU_ADD, atype);
    j.in1 = (const char*)in1; j.in2 = (const char*)in2; j.in3 = (const char*)in3; j.len = len;
    j.s1 = j.s2 = j.isz_in;
    j.host_out = out; j.out_bytes = j.isz_in;
    return dispatch_fused(j, max_workers);
}

}  // extern "C"
