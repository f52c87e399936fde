// Shared internals of libpnumpy_cuda: status handling, launch accounting, tunables,
// per-device properties.  Not part of the public ABI (see include/pnumpy_cuda.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include <vector>
#include "../../../include/pnumpy_cuda.h"

namespace pncu {

// thread-local error text ---------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define PNCU_CUDA(call)                                                         \
    do {                                                                        \
        cudaError_t _e = (call);                                                \
        if (_e != cudaSuccess) return ::pncu::cuda_fail(_e, #call, __FILE__, __LINE__); \
    } while (0)

// launch accounting -----------------------------------------------------------
extern std::atomic<int64_t> g_launches;

inline int after_launch(const char* what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        cudaGetLastError();
        return cuda_fail(e, what, __FILE__, __LINE__);
    }
    return 0;
}

// tunables --------------------------------------------------------------------
struct Tunables {
    int ctas_per_sm;         // persistent grid = SMs * ctas_per_sm for element-wise kernels
    int reduce_ctas_per_sm;  // same for reduction pass 1
};
extern Tunables g_tune;

// device properties (cached per device) ----------------------------------------
struct DevProps {
    int sm_count;
    int cc_major, cc_minor;
    size_t total_mem;
    char name[128];
    bool valid;
};
const DevProps* dev_props();          // of the current device, NULL on failure
int ensure_init();                    // 0 or PNCU_ERR_NO_DEVICE

// Per-(device, stream) scratch for reduction partials; grows on demand, never shrinks.
// `ticket` (optional): the (device, stream)'s zeroed 32-bit counter for one-launch reductions.
int get_workspace(cudaStream_t stream, size_t bytes, void** out, unsigned** ticket = nullptr, size_t ngtickets = 0,
                  unsigned** gtickets = nullptr);
int reset_ticket(cudaStream_t stream);

// ---- allocation registry ---------------------------------------------------------------------------
// Every block this library hands out (pncu_malloc, pncu_malloc_host, pncu_host_register, nte_alloc_pinned,
// nte_alloc_managed) is recorded by base address, so that the dispatcher can tell what kind of memory an
// operand lives in with one ordered-map lookup under a shared lock (~30 ns) instead of a driver call per
// operand per call.  Pointers that are not in the registry are classified by cudaPointerGetAttributes.
enum MemKind { MK_UNKNOWN = -1, MK_PAGEABLE = 0, MK_PINNED = 1, MK_DEVICE = 2, MK_MANAGED = 3 };
struct ResidencyNote {      // "the managed range [lo, lo + bytes) was prefetched to ..."
    const char* lo;
    size_t bytes;
    int where;              // GPU index, or 100 + G: sharded over G GPUs in pieces of `piece` bytes
    size_t piece;
};
struct AllocInfo {
    const char* base;
    size_t bytes;
    int kind;               // MemKind
    int device;             // owning device (MK_DEVICE), -1 otherwise
    std::vector<ResidencyNote>* notes;  // managed blocks only; owned by the registry, guarded by the dispatcher's lock
};
void registry_add(const void* base, size_t bytes, int kind, int device);
bool registry_remove(const void* base, size_t* bytes = nullptr);
bool registry_find(const void* p, AllocInfo* out);   // the block containing p

inline cudaStream_t as_stream(pncu_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// grid for a persistent grid-stride kernel over `tiles` tiles
inline unsigned grid_for(int64_t tiles, int ctas_per_sm) {
    const DevProps* p = dev_props();
    int64_t cap = (int64_t)(p ? p->sm_count : 148) * ctas_per_sm;
    int64_t g = tiles < cap ? tiles : cap;
    if (g < 1) g = 1;
    return (unsigned)g;
}

}  // namespace pncu
