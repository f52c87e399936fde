// Experiment (not part of the product): does staging the streams through shared memory with the
// bulk-copy engine (cp.async.bulk + mbarrier, SASS UBLKCP) beat the register path (256-bit
// ld.global / st.global, full grid) that libpnumpy_cuda uses for its HBM-bound kernels?
//   add  (2R+1W)  variant L : bulk loads -> smem ring, results stored from registers
//                 variant LS: bulk loads -> smem ring, results -> smem -> bulk store
//   sum  (1R)     bulk loads -> smem ring, fold from smem
// Persistent grid, SMs x (CTAs that fit in 227 KiB of shared memory), 256 threads per CTA.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/tma_stream tools/tma_stream.cu
//   gpurun -- ./tools/tma_stream [log2_elems]
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}

// MODE 0 = add, bulk loads + register stores; 1 = add, bulk loads + bulk stores; 2 = sum
template <int MODE, int VPT, int STAGES>
__global__ void __launch_bounds__(256) tma_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ o,
                                                  long ntiles, float* sink) {
    constexpr int TILE_F = 256 * 4 * VPT;  // floats per tile per operand
    constexpr uint32_t TILE_B = TILE_F * 4;
    constexpr int NOP = MODE == 2 ? 1 : 2;
    extern __shared__ __align__(128) unsigned char smem[];
    float* ring = reinterpret_cast<float*>(smem);                         // [STAGES][NOP][TILE_F]
    float* obuf = ring + (size_t)STAGES * NOP * TILE_F;                   // [3][TILE_F] (MODE 1)
    uint64_t* full = reinterpret_cast<uint64_t*>(obuf + (MODE == 1 ? 3 * TILE_F : 0));
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long first = blockIdx.x, step = gridDim.x;
    auto issue = [&](long t, int s) {
        mbar_expect_tx(&full[s], TILE_B * NOP);
        bulk_g2s(ring + ((size_t)s * NOP + 0) * TILE_F, a + t * TILE_F, TILE_B, &full[s]);
        if (NOP == 2) bulk_g2s(ring + ((size_t)s * NOP + 1) * TILE_F, b + t * TILE_F, TILE_B, &full[s]);
    };
    if (tid == 0)
        for (int s = 0; s < STAGES; ++s)
            if (first + s * step < ntiles) issue(first + s * step, s);
    float acc = 0.f;
    long k = 0;
    for (long t = first; t < ntiles; t += step, ++k) {
        const int s = (int)(k % STAGES);
        mbar_wait(&full[s], (uint32_t)((k / STAGES) & 1));
        const float4* xa = reinterpret_cast<const float4*>(ring + ((size_t)s * NOP + 0) * TILE_F);
        const float4* xb = reinterpret_cast<const float4*>(ring + ((size_t)s * NOP + 1) * TILE_F);
        float4 x[VPT];
#pragma unroll
        for (int v = 0; v < VPT; ++v) x[v] = xa[v * 256 + tid];
        if (MODE != 2) {
#pragma unroll
            for (int v = 0; v < VPT; ++v) {
                float4 y = xb[v * 256 + tid];
                x[v].x += y.x; x[v].y += y.y; x[v].z += y.z; x[v].w += y.w;
            }
        }
        if (MODE == 0) {
            float4* og = reinterpret_cast<float4*>(o + t * TILE_F);
#pragma unroll
            for (int v = 0; v < VPT; ++v) og[v * 256 + tid] = x[v];
        } else if (MODE == 1) {
            float4* os = reinterpret_cast<float4*>(obuf + (size_t)(k % 3) * TILE_F);
#pragma unroll
            for (int v = 0; v < VPT; ++v) os[v * 256 + tid] = x[v];
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        } else {
#pragma unroll
            for (int v = 0; v < VPT; ++v) acc += (x[v].x + x[v].y) + (x[v].z + x[v].w);
        }
        __syncthreads();  // everyone is done with ring stage s (and, MODE 1, has written obuf[k%3])
        if (tid == 0) {
            if (MODE == 1) {
                bulk_s2g(o + t * TILE_F, obuf + (size_t)(k % 3) * TILE_F, TILE_B);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            }
            const long tn = t + (long)STAGES * step;
            if (tn < ntiles) issue(tn, s);
        }
    }
    if (MODE == 1 && tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (MODE == 2 && acc == 12345.678f) *sink = acc;
}

static float *A, *B, *O, *SINK;
static long N;
static int SMS;

template <int MODE, int VPT, int STAGES>
void run(int max_cps) {
    constexpr int TILE_F = 256 * 4 * VPT;
    constexpr int NOP = MODE == 2 ? 1 : 2;
    const size_t smem = (size_t)STAGES * NOP * TILE_F * 4 + (MODE == 1 ? 3 * TILE_F * 4 : 0) + STAGES * 8 + 128;
    if (smem > 227 * 1024) return;
    auto kern = tma_kernel<MODE, VPT, STAGES>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int cps = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cps, kern, 256, smem));
    cps = std::min(cps, max_cps);
    if (cps < 1) return;
    const long ntiles = N / TILE_F;
    const long grid = std::min<long>(ntiles, (long)SMS * cps);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) kern<<<(unsigned)grid, 256, smem>>>(A, B, O, ntiles, SINK);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    float best = 1e30f, tot = 0;
    const int reps = 10;
    for (int i = 0; i < reps; ++i) {
        CK(cudaEventRecord(e0));
        kern<<<(unsigned)grid, 256, smem>>>(A, B, O, ntiles, SINK);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        best = std::min(best, ms); tot += ms;
    }
    const double bytes = (MODE == 2 ? 4.0 : 12.0) * N;
    const char* names[] = {"add  bulk-load/reg-store ", "add  bulk-load/bulk-store", "sum  bulk-load           "};
    printf("%s tile=%2d KiB stages=%d ctas/sm=%d smem=%3zu KiB  avg %.4f ms  %8.1f GB/s (best %8.1f)\n", names[MODE], TILE_F * 4 / 1024,
           STAGES, cps, smem / 1024, tot / reps, bytes / (tot / reps * 1e-3) / 1e9, bytes / (best * 1e-3) / 1e9);
    fflush(stdout);
    CK(cudaEventDestroy(e0)); CK(cudaEventDestroy(e1));
}

template <int MODE>
void sweep() {
    for (int cps : {8, 2, 1}) {
        run<MODE, 1, 2>(cps); run<MODE, 1, 4>(cps); run<MODE, 1, 8>(cps);
        run<MODE, 2, 2>(cps); run<MODE, 2, 3>(cps); run<MODE, 2, 4>(cps);
        run<MODE, 4, 2>(cps); run<MODE, 4, 3>(cps); run<MODE, 4, 4>(cps);
        run<MODE, 8, 2>(cps); run<MODE, 8, 3>(cps);
    }
}

int main(int argc, char** argv) {
    int lg = argc > 1 ? atoi(argv[1]) : 28;
    N = 1L << lg;
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    SMS = p.multiProcessorCount;
    printf("device %s, %d SMs, N = 2^%d floats\n", p.name, SMS, lg);
    CK(cudaMalloc(&A, N * 4)); CK(cudaMalloc(&B, N * 4)); CK(cudaMalloc(&O, N * 4)); CK(cudaMalloc(&SINK, 4));
    CK(cudaMemset(A, 0, N * 4)); CK(cudaMemset(B, 0, N * 4)); CK(cudaMemset(O, 0, N * 4));
    // correctness of the add variants on a small pattern
    {
        const long n = 1 << 22;
        float* h = (float*)malloc(n * 4);
        for (long i = 0; i < n; ++i) h[i] = (float)(i % 1000);
        CK(cudaMemcpy(A, h, n * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(B, h, n * 4, cudaMemcpyHostToDevice));
        for (int mode = 0; mode < 2; ++mode) {
            CK(cudaMemset(O, 0, n * 4));
            const size_t smem = 3 * 2 * 2048 * 4 + 3 * 2048 * 4 + 3 * 8 + 128;
            if (mode == 0) { CK(cudaFuncSetAttribute(tma_kernel<0, 2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); tma_kernel<0, 2, 3><<<37, 256, smem>>>(A, B, O, n / 2048, SINK); }
            else { CK(cudaFuncSetAttribute(tma_kernel<1, 2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); tma_kernel<1, 2, 3><<<37, 256, smem>>>(A, B, O, n / 2048, SINK); }
            CK(cudaDeviceSynchronize());
            float* r = (float*)malloc(n * 4);
            CK(cudaMemcpy(r, O, n * 4, cudaMemcpyDeviceToHost));
            long bad = 0;
            for (long i = 0; i < n; ++i) bad += r[i] != 2 * h[i];
            printf("check mode %d: %ld mismatches\n", mode, bad);
            free(r);
        }
        free(h);
        CK(cudaMemset(A, 0, N * 4)); CK(cudaMemset(B, 0, N * 4));
    }
    sweep<0>();
    sweep<1>();
    sweep<2>();
    return 0;
}
