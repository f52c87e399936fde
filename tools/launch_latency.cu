// How fast can the host learn that a kernel has finished?  The fixed cost of a resident call through the hooks is
// launch + completion detection (~13 us of a 15 us np.add on 1 M elements), so this measures, on one GPU, the wall
// time from just before the launch to the moment the host knows the kernel is done, for a kernel of ~0, ~20 and
// ~60 us, with five ways of waiting:
//   sync        cudaStreamSynchronize
//   query       spin on cudaStreamQuery
//   event       cudaEventRecord + spin on cudaEventQuery
//   flag2       a second 1-thread kernel on the same stream stores a sequence number into MAPPED host memory; host spins on it
//   flag2-pdl   the same, the flag kernel launched with programmatic stream serialization (it is resident and waiting
//               in griddepcontrol.wait when the first kernel ends)
//   ticket      the kernel's last CTA (atomic ticket) stores the flag itself
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/launch_latency tools/launch_latency.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>
#include <algorithm>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
static double now_us() { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec * 1e6 + t.tv_nsec * 1e-3; }

__global__ void work(const float4* a, float4* o, long n, unsigned* ticket, volatile unsigned long long* flag, unsigned long long v) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { float4 x = a[i]; x.x += 1.f; o[i] = x; }
    if (ticket) {
        __shared__ int last;
        __syncthreads();
        if (threadIdx.x == 0) { __threadfence(); last = atomicAdd(ticket, 1u) == gridDim.x - 1; }
        __syncthreads();
        if (last && threadIdx.x == 0) { *ticket = 0; __threadfence_system(); *flag = v; }
    }
}
__global__ void set_flag(volatile unsigned long long* flag, unsigned long long v) { *flag = v; }
__global__ void set_flag_pdl(volatile unsigned long long* flag, unsigned long long v) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    *flag = v;
}

int main(int argc, char** argv) {
    const bool managed = argc > 1 && argv[1][0] == 'm';   // `launch_latency m`: the buffers are cudaMallocManaged, prefetched
    CK(cudaSetDevice(0));
    cudaStream_t st;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    cudaEvent_t ev;
    CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    unsigned long long* flag;
    CK(cudaHostAlloc((void**)&flag, 64, cudaHostAllocMapped));
    *flag = 0;
    unsigned* ticket;
    CK(cudaMalloc((void**)&ticket, 4));
    CK(cudaMemset(ticket, 0, 4));
    const long sizes[3] = {1 << 10, 1 << 23, 1 << 25};   // float4 elements: ~0, ~35, ~140 MB of traffic
    float4 *a, *o;
    if (managed) {
        CK(cudaMallocManaged((void**)&a, (size_t)sizes[2] * 16));
        CK(cudaMallocManaged((void**)&o, (size_t)sizes[2] * 16));
        CK(cudaMemPrefetchAsync(a, (size_t)sizes[2] * 16, 0, st));
        CK(cudaMemPrefetchAsync(o, (size_t)sizes[2] * 16, 0, st));
        CK(cudaStreamSynchronize(st));
        printf("buffers: cudaMallocManaged, prefetched to GPU 0\n");
    } else {
        CK(cudaMalloc((void**)&a, (size_t)sizes[2] * 16));
        CK(cudaMalloc((void**)&o, (size_t)sizes[2] * 16));
        printf("buffers: cudaMalloc\n");
    }
    CK(cudaMemset(a, 0, (size_t)sizes[2] * 16));
    unsigned long long seq = 0;
    const char* names[6] = {"sync", "query", "event", "flag2", "flag2-pdl", "ticket"};
    for (long n : sizes) {
        const unsigned grid = (unsigned)((n + 255) / 256);
        // kernel time alone
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        for (int i = 0; i < 5; ++i) work<<<grid, 256, 0, st>>>(a, o, n, nullptr, nullptr, 0);
        CK(cudaEventRecord(e0, st));
        for (int i = 0; i < 20; ++i) work<<<grid, 256, 0, st>>>(a, o, n, nullptr, nullptr, 0);
        CK(cudaEventRecord(e1, st));
        CK(cudaStreamSynchronize(st));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("n = %ld float4 (%.0f MB moved): kernel alone %.1f us (back to back)\n", n, n * 32e-6, ms * 1e3 / 20);
        for (int mode = 0; mode < 6; ++mode) {
            std::vector<double> ts, tl;
            for (int it = 0; it < 300; ++it) {
                const double t0 = now_us();
                double t1;
                ++seq;
                if (mode == 5) work<<<grid, 256, 0, st>>>(a, o, n, ticket, flag, seq);
                else work<<<grid, 256, 0, st>>>(a, o, n, nullptr, nullptr, 0);
                if (mode == 0) { t1 = now_us(); CK(cudaStreamSynchronize(st)); }
                else if (mode == 1) { t1 = now_us(); while (cudaStreamQuery(st) == cudaErrorNotReady) {} }
                else if (mode == 2) { CK(cudaEventRecord(ev, st)); t1 = now_us(); while (cudaEventQuery(ev) == cudaErrorNotReady) {} }
                else if (mode == 3) { set_flag<<<1, 1, 0, st>>>(flag, seq); t1 = now_us(); while (*(volatile unsigned long long*)flag != seq) {} }
                else if (mode == 4) {
                    cudaLaunchConfig_t cfg = {};
                    cfg.gridDim = dim3(1); cfg.blockDim = dim3(1); cfg.stream = st;
                    cudaLaunchAttribute at;
                    at.id = cudaLaunchAttributeProgrammaticStreamSerialization;
                    at.val.programmaticStreamSerializationAllowed = 1;
                    cfg.attrs = &at; cfg.numAttrs = 1;
                    CK(cudaLaunchKernelEx(&cfg, set_flag_pdl, (volatile unsigned long long*)flag, seq));
                    t1 = now_us();
                    while (*(volatile unsigned long long*)flag != seq) {}
                } else { t1 = now_us(); while (*(volatile unsigned long long*)flag != seq) {} }
                const double t2 = now_us();
                if (it >= 50) { ts.push_back(t2 - t0); tl.push_back(t1 - t0); }
            }
            CK(cudaStreamSynchronize(st));
            std::sort(ts.begin(), ts.end()); std::sort(tl.begin(), tl.end());
            printf("  %-10s total median %6.1f us  best %6.1f   (launch calls: median %4.1f us)\n", names[mode], ts[ts.size() / 2], ts[0], tl[tl.size() / 2]);
        }
    }
    return 0;
}
