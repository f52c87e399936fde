// Shared internals of libpnumpy_cuda: status handling, launch accounting, tunables,
// per-device properties.  Not part of the public ABI (see include/pnumpy_cuda.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
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
int get_workspace(cudaStream_t stream, size_t bytes, void** out);

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
