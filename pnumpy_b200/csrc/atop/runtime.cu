// Library lifetime, error text, tunables, reduction scratch and the thin cudart wrappers of
// include/pnumpy_cuda.h.  replaces: atop_init() (src/atop/atop.cpp:68-104) and the CPU-identity
// half of CMathWorker (src/atop/threads.h:679-701, threads.cpp:616-656).
#include <stdarg.h>
#include <string.h>
#include <map>
#include <mutex>
#include <shared_mutex>
#include <utility>
#include "common.cuh"

namespace pncu {

std::atomic<int64_t> g_launches{0};
Tunables g_tune = {/*ctas_per_sm=*/8, /*reduce_ctas_per_sm=*/8};

static thread_local char t_error[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_error, sizeof(t_error), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    const char* base = strrchr(file, '/');
    set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), base ? base + 1 : file, line, what);
    return (int)e;
}

static std::mutex g_mu;
static int g_ndev = -1;  // -1 = not probed
static const int kMaxDev = 64;
static DevProps g_props[kMaxDev];

int ensure_init() {
    if (g_ndev > 0) return 0;
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_ndev > 0) return 0;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        cudaGetLastError();
        set_error("no usable CUDA device (%s); libpnumpy_cuda has no CPU fallback",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
        g_ndev = 0;
        return PNCU_ERR_NO_DEVICE;
    }
    if (n > kMaxDev) n = kMaxDev;
    for (int d = 0; d < n; ++d) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, d) != cudaSuccess) { cudaGetLastError(); g_props[d].valid = false; continue; }
        g_props[d].sm_count = p.multiProcessorCount;
        g_props[d].cc_major = p.major;
        g_props[d].cc_minor = p.minor;
        g_props[d].total_mem = p.totalGlobalMem;
        snprintf(g_props[d].name, sizeof(g_props[d].name), "%.127s", p.name);
        g_props[d].valid = true;
    }
    // the fat binary holds sm_100a SASS only: refuse anything else up front with a clear message
    if (!g_props[0].valid || g_props[0].cc_major != 10) {
        set_error("device 0 is '%s' (sm_%d%d); libpnumpy_cuda is built for sm_100a (B200) only",
                  g_props[0].valid ? g_props[0].name : "?", g_props[0].cc_major, g_props[0].cc_minor);
        g_ndev = 0;
        return PNCU_ERR_NO_DEVICE;
    }
    g_ndev = n;
    return 0;
}

const DevProps* dev_props() {
    int d = 0;
    if (g_ndev <= 0 || cudaGetDevice(&d) != cudaSuccess || d >= g_ndev) return NULL;
    return g_props[d].valid ? &g_props[d] : NULL;
}

// ---- reduction scratch, keyed by (device, stream) --------------------------------------------
// `ticket`: one 32-bit counter per (device, stream), zero between launches (the last CTA of a one-launch
// reduction re-arms it); lives in its own allocation so that growing the scratch never disturbs it.
struct Scratch { void* ptr; size_t bytes; unsigned* ticket; unsigned* gtickets; size_t ngtickets; };
static std::map<std::pair<int, cudaStream_t>, Scratch> g_scratch;

int get_workspace(cudaStream_t stream, size_t bytes, void** out, unsigned** ticket, size_t ngtickets, unsigned** gtickets) {
    int d = 0;
    PNCU_CUDA(cudaGetDevice(&d));
    std::lock_guard<std::mutex> lk(g_mu);
    Scratch& s = g_scratch[std::make_pair(d, stream)];
    if (ticket && !s.ticket) {
        void* t = NULL;
        cudaError_t e = cudaMalloc(&t, 256);
        if (e != cudaSuccess) { cudaGetLastError(); set_error("cudaMalloc(256) for a reduction ticket failed: %s", cudaGetErrorString(e)); return PNCU_ERR_NOMEM; }
        PNCU_CUDA(cudaMemset(t, 0, 256));   // synchronous: visible to whatever stream uses it next
        s.ticket = (unsigned*)t;
    }
    if (s.bytes < bytes) {
        size_t want = bytes < (1u << 20) ? (1u << 20) : bytes * 2;
        void* p = NULL;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { cudaGetLastError(); set_error("cudaMalloc(%zu) for reduction scratch failed: %s", want, cudaGetErrorString(e)); return PNCU_ERR_NOMEM; }
        if (s.ptr) {
            // the old buffer may still be read by work queued on this stream
            cudaStreamSynchronize(stream);
            cudaFree(s.ptr);
        }
        s.ptr = p;
        s.bytes = want;
    }
    if (gtickets && s.ngtickets < ngtickets) {
        // group tickets (one 32-bit counter per 64 slots): zero between launches, so a larger array starts zeroed
        size_t want = ngtickets < 4096 ? 4096 : ngtickets * 2;
        void* p = NULL;
        cudaError_t e = cudaMalloc(&p, want * sizeof(unsigned));
        if (e != cudaSuccess) { cudaGetLastError(); set_error("cudaMalloc(%zu) for group tickets failed: %s", want * sizeof(unsigned), cudaGetErrorString(e)); return PNCU_ERR_NOMEM; }
        PNCU_CUDA(cudaMemset(p, 0, want * sizeof(unsigned)));
        if (s.gtickets) {
            cudaStreamSynchronize(stream);   // a kernel queued on this stream may still count in the old array
            cudaFree(s.gtickets);
        }
        s.gtickets = (unsigned*)p;
        s.ngtickets = want;
    }
    if (out) *out = s.ptr;
    if (ticket) *ticket = s.ticket;
    if (gtickets) *gtickets = s.gtickets;
    return 0;
}

// after a failed / timed-out launch the counter may be left non-zero
int reset_ticket(cudaStream_t stream) {
    int d = 0;
    PNCU_CUDA(cudaGetDevice(&d));
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_scratch.find(std::make_pair(d, stream));
    if (it != g_scratch.end() && it->second.ticket) PNCU_CUDA(cudaMemsetAsync(it->second.ticket, 0, 256, stream));
    if (it != g_scratch.end() && it->second.gtickets)
        PNCU_CUDA(cudaMemsetAsync(it->second.gtickets, 0, it->second.ngtickets * sizeof(unsigned), stream));
    return 0;
}

// ---- allocation registry (common.cuh) -------------------------------------------------------------------
static std::shared_mutex g_reg_mu;
static std::map<uintptr_t, AllocInfo> g_registry;

void registry_add(const void* base, size_t bytes, int kind, int device) {
    if (!base) return;
    AllocInfo a;
    a.base = (const char*)base;
    a.bytes = bytes ? bytes : 1;
    a.kind = kind;
    a.device = device;
    a.notes = kind == MK_MANAGED ? new std::vector<ResidencyNote>() : nullptr;
    std::unique_lock<std::shared_mutex> lk(g_reg_mu);
    auto it = g_registry.find((uintptr_t)base);
    if (it != g_registry.end()) { delete it->second.notes; g_registry.erase(it); }
    g_registry[(uintptr_t)base] = a;
}

bool registry_remove(const void* base, size_t* bytes) {
    std::unique_lock<std::shared_mutex> lk(g_reg_mu);
    auto it = g_registry.find((uintptr_t)base);
    if (it == g_registry.end()) return false;
    if (bytes) *bytes = it->second.bytes;
    delete it->second.notes;
    g_registry.erase(it);
    return true;
}

bool registry_find(const void* p, AllocInfo* out) {
    std::shared_lock<std::shared_mutex> lk(g_reg_mu);
    if (g_registry.empty()) return false;
    auto it = g_registry.upper_bound((uintptr_t)p);
    if (it == g_registry.begin()) return false;
    --it;
    if ((uintptr_t)p >= it->first + it->second.bytes) return false;
    if (out) *out = it->second;
    return true;
}

}  // namespace pncu

using namespace pncu;

extern "C" {

int pncu_init(int* device_count) {
    int rc = ensure_init();
    if (device_count) *device_count = g_ndev > 0 ? g_ndev : 0;
    return rc;
}

int pncu_shutdown(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    for (auto& kv : g_scratch) {
        if (kv.second.ptr || kv.second.ticket) {
            cudaSetDevice(kv.first.first);
            if (kv.second.ptr) cudaFree(kv.second.ptr);
            if (kv.second.ticket) cudaFree(kv.second.ticket);
        }
    }
    g_scratch.clear();
    return 0;
}

int pncu_device_count(void) { return ensure_init() ? 0 : g_ndev; }

int pncu_device_info(int device, char* buf, size_t buflen) {
    int rc = ensure_init();
    if (rc) return rc;
    if (!buf || buflen == 0 || device < 0 || device >= g_ndev || !g_props[device].valid) {
        set_error("bad device index %d", device);
        return PNCU_ERR_ARGUMENT;
    }
    const DevProps& p = g_props[device];
    snprintf(buf, buflen, "%s sm_%d%d %d SMs %.1f GiB HBM", p.name, p.cc_major, p.cc_minor, p.sm_count,
             (double)p.total_mem / (1024.0 * 1024.0 * 1024.0));
    return 0;
}

const char* pncu_last_error(void) { return t_error; }
int64_t pncu_launch_count(void) { return g_launches.load(); }

int pncu_set_tunable(const char* name, int value) {
    if (!name || value < 1 || value > 64) return -1;
    int* slot = NULL;
    if (!strcmp(name, "ctas_per_sm")) slot = &g_tune.ctas_per_sm;
    else if (!strcmp(name, "reduce_ctas_per_sm")) slot = &g_tune.reduce_ctas_per_sm;
    if (!slot) return -1;
    int prev = *slot;
    *slot = value;
    return prev;
}

int pncu_itemsize(int atype) {
    // src/pnumpy/_pnumpy.cpp:43-54, restricted to the types that have kernels
    switch (atype) {
        case PNCU_BOOL: case PNCU_INT8: case PNCU_UINT8: return 1;
        case PNCU_INT16: case PNCU_UINT16: return 2;
        case PNCU_INT32: case PNCU_UINT32: case PNCU_FLOAT: return 4;
        case PNCU_INT64: case PNCU_UINT64: case PNCU_DOUBLE: return 8;
    }
    return 0;
}

int pncu_set_device(int device) { int rc = ensure_init(); if (rc) return rc; PNCU_CUDA(cudaSetDevice(device)); return 0; }
int pncu_get_device(int* device) { int rc = ensure_init(); if (rc) return rc; PNCU_CUDA(cudaGetDevice(device)); return 0; }
int pncu_malloc(void** ptr, size_t bytes) {
    int rc = ensure_init(); if (rc) return rc;
    PNCU_CUDA(cudaMalloc(ptr, bytes));
    int d = 0;
    cudaGetDevice(&d);
    registry_add(*ptr, bytes, MK_DEVICE, d);
    return 0;
}
int pncu_free(void* ptr) { registry_remove(ptr); PNCU_CUDA(cudaFree(ptr)); return 0; }
int pncu_malloc_host(void** ptr, size_t bytes) {
    int rc = ensure_init(); if (rc) return rc;
    PNCU_CUDA(cudaHostAlloc(ptr, bytes, cudaHostAllocPortable | cudaHostAllocMapped));
    registry_add(*ptr, bytes, MK_PINNED, -1);
    return 0;
}
int pncu_free_host(void* ptr) { registry_remove(ptr); PNCU_CUDA(cudaFreeHost(ptr)); return 0; }
int pncu_host_register(void* ptr, size_t bytes) {
    int rc = ensure_init(); if (rc) return rc;
    PNCU_CUDA(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped));
    registry_add(ptr, bytes, MK_PINNED, -1);
    return 0;
}
int pncu_host_unregister(void* ptr) { registry_remove(ptr); PNCU_CUDA(cudaHostUnregister(ptr)); return 0; }
int pncu_memcpy_h2d(void* dst, const void* src, size_t bytes, pncu_stream_t s) {
    PNCU_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, as_stream(s))); return 0;
}
int pncu_memcpy_d2h(void* dst, const void* src, size_t bytes, pncu_stream_t s) {
    PNCU_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, as_stream(s))); return 0;
}
int pncu_memcpy_d2d(void* dst, const void* src, size_t bytes, pncu_stream_t s) {
    PNCU_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, as_stream(s))); return 0;
}
int pncu_memset(void* dst, int value, size_t bytes, pncu_stream_t s) {
    PNCU_CUDA(cudaMemsetAsync(dst, value, bytes, as_stream(s))); return 0;
}
int pncu_stream_create(pncu_stream_t* s) {
    int rc = ensure_init(); if (rc) return rc;
    cudaStream_t st;
    PNCU_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    *s = (pncu_stream_t)st;
    return 0;
}
int pncu_stream_destroy(pncu_stream_t s) { PNCU_CUDA(cudaStreamDestroy(as_stream(s))); return 0; }
int pncu_stream_sync(pncu_stream_t s) { PNCU_CUDA(cudaStreamSynchronize(as_stream(s))); return 0; }
int pncu_device_sync(void) { PNCU_CUDA(cudaDeviceSynchronize()); return 0; }
int pncu_event_create(pncu_event_t* ev) {
    int rc = ensure_init(); if (rc) return rc;
    cudaEvent_t e;
    PNCU_CUDA(cudaEventCreate(&e));
    *ev = (pncu_event_t)e;
    return 0;
}
int pncu_event_destroy(pncu_event_t ev) { PNCU_CUDA(cudaEventDestroy((cudaEvent_t)ev)); return 0; }
int pncu_event_record(pncu_event_t ev, pncu_stream_t s) { PNCU_CUDA(cudaEventRecord((cudaEvent_t)ev, as_stream(s))); return 0; }
int pncu_event_sync(pncu_event_t ev) { PNCU_CUDA(cudaEventSynchronize((cudaEvent_t)ev)); return 0; }
int pncu_event_elapsed_ms(pncu_event_t a, pncu_event_t b, float* ms) {
    PNCU_CUDA(cudaEventElapsedTime(ms, (cudaEvent_t)a, (cudaEvent_t)b)); return 0;
}
int pncu_pointer_kind(const void* ptr, int* device) {
    int rc = ensure_init(); if (rc) return -1;
    AllocInfo ai;
    if (registry_find(ptr, &ai)) {
        if (device) *device = ai.device;
        return ai.kind;
    }
    cudaPointerAttributes at;
    cudaError_t e = cudaPointerGetAttributes(&at, ptr);
    if (e != cudaSuccess) { cudaGetLastError(); return 0; }
    if (device) *device = at.device;
    switch (at.type) {
        case cudaMemoryTypeHost: return 1;
        case cudaMemoryTypeDevice: return 2;
        case cudaMemoryTypeManaged: return 3;
        default: return 0;
    }
}

// ---- peer memory plumbing for the fused exchange (ops_reduce.cu) -----------------------------------------
int pncu_enable_peer_access(int peer_device) {
    int rc = ensure_init(); if (rc) return rc;
    int cur = 0;
    PNCU_CUDA(cudaGetDevice(&cur));
    if (cur == peer_device) return 0;
    int can = 0;
    PNCU_CUDA(cudaDeviceCanAccessPeer(&can, cur, peer_device));
    if (!can) { set_error("device %d cannot access device %d directly", cur, peer_device); return PNCU_ERR_UNSUPPORTED; }
    cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return 0; }
    if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceEnablePeerAccess", __FILE__, __LINE__);
    return 0;
}
int pncu_ipc_export(void* dev_ptr, void* handle64) {
    int rc = ensure_init(); if (rc) return rc;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    if (!dev_ptr || !handle64) { set_error("null argument"); return PNCU_ERR_ARGUMENT; }
    cudaIpcMemHandle_t h;
    PNCU_CUDA(cudaIpcGetMemHandle(&h, dev_ptr));
    memcpy(handle64, &h, 64);
    return 0;
}
int pncu_ipc_import(const void* handle64, void** dev_ptr) {
    int rc = ensure_init(); if (rc) return rc;
    if (!dev_ptr || !handle64) { set_error("null argument"); return PNCU_ERR_ARGUMENT; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    PNCU_CUDA(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}
int pncu_ipc_close(void* dev_ptr) {
    int rc = ensure_init(); if (rc) return rc;
    PNCU_CUDA(cudaIpcCloseMemHandle(dev_ptr));
    return 0;
}

}  // extern "C"
