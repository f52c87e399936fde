// Tuning microbenchmark for the HBM-streaming kernel shapes used by libpnumpy_cuda (not part of
// the product).  Sweeps vector width (128/256-bit), unroll, block size, cache policy and grid
// policy for add (2R+1W), copy (1R+1W) and a read-only sum, and prints achieved GB/s.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/stream_tune tools/stream_tune.cu
//   gpurun -- ./tools/stream_tune [log2_elems]
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <string>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template <int VB> struct V;
template <> struct V<16> { float v[4]; };
template <> struct V<32> { float v[8]; };

template <int VB, int POL> __device__ __forceinline__ V<VB> ld(const float* p) {
    V<VB> r;
    if constexpr (VB == 16) {
        if constexpr (POL == 0) asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]) : "l"(p));
        if constexpr (POL == 1) asm volatile("ld.global.cs.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]) : "l"(p));
        if constexpr (POL == 2) asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]) : "l"(p));
    } else {
        if constexpr (POL == 0) asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7]) : "l"(p));
        if constexpr (POL == 1) asm volatile("ld.global.cs.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7]) : "l"(p));
        if constexpr (POL == 2) asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7]) : "l"(p));
    }
    return r;
}
template <int VB, int POL> __device__ __forceinline__ void st(float* p, const V<VB>& r) {
    if constexpr (VB == 16) {
        if constexpr (POL == 1) asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(r.v[0]), "f"(r.v[1]), "f"(r.v[2]), "f"(r.v[3]) : "memory");
        else asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(r.v[0]), "f"(r.v[1]), "f"(r.v[2]), "f"(r.v[3]) : "memory");
    } else {
        if constexpr (POL == 1) asm volatile("st.global.cs.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(r.v[0]), "f"(r.v[1]), "f"(r.v[2]), "f"(r.v[3]), "f"(r.v[4]), "f"(r.v[5]), "f"(r.v[6]), "f"(r.v[7]) : "memory");
        else asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(r.v[0]), "f"(r.v[1]), "f"(r.v[2]), "f"(r.v[3]), "f"(r.v[4]), "f"(r.v[5]), "f"(r.v[6]), "f"(r.v[7]) : "memory");
    }
}

// MODE 0 = add (2R 1W), 1 = copy (1R 1W), 2 = read-only sum (1R)
template <int MODE, int VB, int UNROLL, int BLOCK, int POL>
__global__ void __launch_bounds__(BLOCK) stream_kernel(const float* a, const float* b, float* o, long nvec, float* sink) {
    constexpr int E = VB / 4;
    const long tile = (long)BLOCK * UNROLL;
    const long ntiles = nvec / tile;  // sizes are powers of two: no ragged tile
    float acc = 0.f;
    for (long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const long base = t * tile + threadIdx.x;
        V<VB> x[UNROLL], y[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            x[u] = ld<VB, POL>(a + (base + (long)u * BLOCK) * E);
            if (MODE == 0) y[u] = ld<VB, POL>(b + (base + (long)u * BLOCK) * E);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            if (MODE == 0) {
#pragma unroll
                for (int e = 0; e < E; ++e) x[u].v[e] += y[u].v[e];
            }
            if (MODE == 2) {
#pragma unroll
                for (int e = 0; e < E; ++e) acc += x[u].v[e];
            } else {
                st<VB, POL>(o + (base + (long)u * BLOCK) * E, x[u]);
            }
        }
    }
    if (MODE == 2 && acc == 12345.678f) *sink = acc;
}

struct Result { const char* mode; int vb, unroll, block, pol, cps; double gbs; };
static std::vector<Result> results;
static float *A, *B, *O, *SINK;
static long N;
static int SMS;

template <int MODE, int VB, int UNROLL, int BLOCK, int POL>
void run(int cps) {
    const long nvec = N / (VB / 4);
    const long ntiles = nvec / ((long)BLOCK * UNROLL);
    long grid = cps == 0 ? ntiles : std::min<long>(ntiles, (long)SMS * cps);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) stream_kernel<MODE, VB, UNROLL, BLOCK, POL><<<(unsigned)grid, BLOCK>>>(A, B, O, nvec, SINK);
    CK(cudaGetLastError());
    float best = 1e30f, tot = 0;
    const int reps = 10;
    for (int i = 0; i < reps; ++i) {
        CK(cudaEventRecord(e0));
        stream_kernel<MODE, VB, UNROLL, BLOCK, POL><<<(unsigned)grid, BLOCK>>>(A, B, O, nvec, SINK);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        best = std::min(best, ms); tot += ms;
    }
    const double bytes = (MODE == 0 ? 12.0 : MODE == 1 ? 8.0 : 4.0) * N;
    const char* names[] = {"add", "copy", "sum"};
    results.push_back({names[MODE], VB, UNROLL, BLOCK, POL, cps, bytes / (tot / reps * 1e-3) / 1e9});
    printf("%-4s vb=%2d unroll=%d block=%4d pol=%d ctas/sm=%2d  avg %.4f ms  best %.4f ms  %8.1f GB/s (best %8.1f)\n",
           names[MODE], VB, UNROLL, BLOCK, POL, cps, tot / reps, best, bytes / (tot / reps * 1e-3) / 1e9, bytes / (best * 1e-3) / 1e9);
    CK(cudaEventDestroy(e0)); CK(cudaEventDestroy(e1));
}

template <int MODE, int VB, int UNROLL, int BLOCK>
void run_pol() {
    const int threads_cps[] = {0, 2048 / BLOCK, 1024 / BLOCK, 4096 / BLOCK};
    for (int cps : threads_cps) {
        if (cps == 0 && UNROLL * BLOCK < 512) continue;
        run<MODE, VB, UNROLL, BLOCK, 1>(cps);
    }
    run<MODE, VB, UNROLL, BLOCK, 0>(2048 / BLOCK);
    run<MODE, VB, UNROLL, BLOCK, 2>(2048 / BLOCK);
    run<MODE, VB, UNROLL, BLOCK, 0>(0);
}
template <int MODE, int VB, int UNROLL>
void run_block() {
    run_pol<MODE, VB, UNROLL, 128>();
    run_pol<MODE, VB, UNROLL, 256>();
    run_pol<MODE, VB, UNROLL, 512>();
    run_pol<MODE, VB, UNROLL, 1024>();
}
template <int MODE>
void run_mode() {
    run_block<MODE, 16, 1>(); run_block<MODE, 16, 2>(); run_block<MODE, 16, 4>(); run_block<MODE, 16, 8>();
    run_block<MODE, 32, 1>(); run_block<MODE, 32, 2>(); run_block<MODE, 32, 4>();
}

int main(int argc, char** argv) {
    int lg = argc > 1 ? atoi(argv[1]) : 28;
    N = 1L << lg;
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    SMS = p.multiProcessorCount;
    printf("device %s, %d SMs, N = 2^%d floats\n", p.name, SMS, lg);
    CK(cudaMalloc(&A, N * 4)); CK(cudaMalloc(&B, N * 4)); CK(cudaMalloc(&O, N * 4)); CK(cudaMalloc(&SINK, 4));
    CK(cudaMemset(A, 0, N * 4)); CK(cudaMemset(B, 0, N * 4)); CK(cudaMemset(O, 0, N * 4));
    // reference: the driver's own D2D copy (what MEASURED_PEAKS.json's torch copy_ uses)
    {
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        for (int i = 0; i < 3; ++i) CK(cudaMemcpyAsync(O, A, N * 4, cudaMemcpyDeviceToDevice));
        float tot = 0, best = 1e30f;
        for (int i = 0; i < 10; ++i) {
            CK(cudaEventRecord(e0)); CK(cudaMemcpyAsync(O, A, N * 4, cudaMemcpyDeviceToDevice)); CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); tot += ms; best = std::min(best, ms);
        }
        printf("cudaMemcpy D2D: avg %.4f ms -> %.1f GB/s (best %.1f)\n", tot / 10, 8.0 * N / (tot / 10 * 1e-3) / 1e9, 8.0 * N / (best * 1e-3) / 1e9);
    }
    run_mode<0>();
    run_mode<1>();
    run_mode<2>();
    for (const char* m : {"add", "copy", "sum"}) {
        std::vector<Result> r;
        for (auto& x : results) if (x.mode == std::string(m)) r.push_back(x);
        std::sort(r.begin(), r.end(), [](const Result& a, const Result& b) { return a.gbs > b.gbs; });
        printf("\nTOP %s:\n", m);
        for (size_t i = 0; i < r.size() && i < 12; ++i)
            printf("  %8.1f GB/s  vb=%d unroll=%d block=%d pol=%d ctas/sm=%d\n", r[i].gbs, r[i].vb, r[i].unroll, r[i].block, r[i].pol, r[i].cps);
    }
    return 0;
}
