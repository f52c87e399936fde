/*
 * pnumpy_nte.h -- C ABI of the nte dispatcher inside libpnumpy_cuda: the per-GPU stream and
 * shard dispatcher that replaces pnumpy's worker-thread 16K-chunk scheduler.
 *
 * What this replaces
 * ------------------
 * `CMathWorker` and friends, src/atop/threads.h:645-1171 + src/atop/threads.cpp:72-191, as the
 * hook layer uses them (src/pnumpy/_pnumpy.cpp:638-996):
 *     THREADER->GetWorkItem(n)            -> nte_* returns NTE_DECLINED below the size threshold,
 *                                            for re-entrant calls and for layouts it cannot stage
 *     THREADER->WorkMain(item, n, maxthr) -> the nte_binary / nte_unary / nte_reduce / nte_argminmax
 *                                            calls below (synchronous: they return when the result
 *                                            is in the caller's buffers, like WorkMain's spin-wait,
 *                                            threads.h:1148-1152)
 *     SetFutexWakeup/GetFutexWakeup/NoThreading -> nte_set_workers / nte_get_workers /
 *                                            nte_set_threading  (same clamping and return values,
 *                                            threads.h:852-872)
 * The reference splits [0, n) into 16384-element chunks dealt round-robin to <= 16 threads.  Here
 * [0, n) is split into `lanes` contiguous shards (boundaries at multiples of the 65536-element
 * reduction block); lane l drives GPU l % G with its own three streams (H2D, compute, D2H) and a
 * ring of pinned staging slabs, so the copy engines and the SMs of every GPU overlap.  Element-wise
 * work needs no exchange between lanes.  Reductions need one: for RESIDENT operands every GPU runs ONE
 * kernel (pncu_*_fused) whose last CTA stores the shard's block partials into GPU 0's table over NVLink,
 * and GPU 0's last CTA folds the table in fixed order and writes the result into mapped host memory;
 * for host operands the lanes' partials are gathered on GPU 0 (peer copies) and folded by one finish
 * kernel -- the "second pass" of src/pnumpy/_pnumpy.cpp:697-699 either way, so the result is
 * independent of the number of lanes and GPUs.
 *
 * Operands are HOST pointers as NumPy hands them to a ufunc inner loop (pageable, or pinned when
 * they came from nte_alloc_pinned / the recycler), or device / managed pointers, which are used in
 * place with no PCIe traffic AT EVERY SIZE (no threshold applies to them); a call may mix the two
 * (resident arrays are used in place, host arrays are staged).  Strides are in bytes, as in the
 * reference's loop types: 0 broadcasts a scalar, the item size is a contiguous array, any other
 * multiple of the item size (negative included) is a strided view whose elements the lanes gather
 * into / scatter from their pinned slabs (the reference's generic strided loop,
 * src/atop/ops_binary.cpp:563-569); resident strided views go to the strided kernels as they are.
 */
#ifndef PNUMPY_NTE_H
#define PNUMPY_NTE_H

#include "pnumpy_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

/* returned by the nte_* ops when the caller should run its previous (NumPy) loop instead:
 * fewer elements than the threshold, dispatcher disabled or busy on another thread, or a layout
 * that cannot be staged (a stride that is not a multiple of the item size, an output that overlaps an
 * input without aliasing it exactly).  Never a failure. */
#define NTE_DECLINED 1

/* Bring up the dispatcher: pncu_init + per-GPU lane state on first use.  Returns 0, or
 * PNCU_ERR_NO_DEVICE (there is no CPU mode).  Idempotent.   replaces: atop_init()'s
 * `new CMathWorker` + StartWorkerThreads (src/atop/atop.cpp:96-101). */
PNCU_API int nte_init(void);
PNCU_API int nte_shutdown(void);
PNCU_API int nte_device_count(void);

/* thread_* controls.  `workers` is the number of EXTRA lanes beside the caller's, exactly like the
 * reference's futex wake count: a call uses min(op_max_workers, workers) + 1 lanes.
 * nte_set_workers clamps to [1, nte_max_workers()] and returns the PREVIOUS value
 * (src/atop/threads.h:852-872). */
PNCU_API int nte_set_workers(int workers);
PNCU_API int nte_get_workers(void);
PNCU_API int nte_max_workers(void);
/* NoThreading flag (src/atop/threads.h:1058): when threading is off a call uses one lane on GPU 0 */
PNCU_API void nte_set_threading(int enabled);
PNCU_API int nte_get_threading(void);

/* Size policy.  The reference threads a call iff n >= 65536 (src/atop/threads.h:199,1058); the
 * analogue here is "worth a PCIe round trip".  Returns the previous value. */
/* Size threshold.  `n` is the base: streaming element-wise ops on pinned operands decline below it; the
 * effective value is x4 when an operand is pageable, /2 for NTE_CLASS_HEAVY ops (transcendentals), x2 for
 * reductions, and it is applied to BYTES (n x 4 bytes of the widest operand; x3 for 1-byte types) -- the measured
 * crossovers against NumPy's own loops (tools/threshold_probe.py, tools/pn_benchmark.py).  A call whose
 * array operands are ALL resident (device / managed memory) always runs on the GPU, whatever its size; a call that
 * mixes host and resident arrays is held to the threshold like a host call (NumPy feeds np.sin(int16_array) through
 * 8192-element cast buffers).  Default 2^19; 0 sends everything to the GPU. */
#define NTE_CLASS_STREAM 0
#define NTE_CLASS_HEAVY 1
#define NTE_CLASS_REDUCE 2
#define NTE_CLASS_CONST 3   /* output independent of the input (integer isnan ...): host operands are always declined */
PNCU_API int64_t nte_set_min_elems(int64_t n);
PNCU_API int64_t nte_get_min_elems(void);
/* bytes of one staging slab (rounded to a multiple of 65536 * 8); returns the previous value */
PNCU_API int64_t nte_set_slab_bytes(int64_t bytes);
/* Fault injection for the error-path tests (also env PNUMPY_CUDA_FAULT_INJECT=n at nte_init): the n-th
 * dispatch from now that would have gone to the GPU returns cudaErrorLaunchFailure (719) without touching
 * the device, exactly as a failed launch would.  0 disarms.  Returns the previous countdown. */
PNCU_API int64_t nte_inject_fault(int64_t nth);
/* 1 in a fork()ed child of a process that had initialised the dispatcher.  CUDA cannot be used there: every
 * nte_* op returns NTE_DECLINED (the caller's NumPy loop runs), the allocators return NULL, the free
 * functions only forget the block.  Start worker processes with the "spawn" method to use the GPU in them. */
PNCU_API int nte_is_forked_child(void);

/* ---- the four dispatch entry points (one per hook-layer dispatcher) --------------------------------
 * `launcher` is what pncu_Get*OpFast returned; item sizes are of one input / output element.
 * max_workers is the per-(ufunc, dtype) MaxThreads of the hook layer (3 or 7,
 * src/pnumpy/_pnumpy.cpp:1373,1544).  Return 0 (done), NTE_DECLINED, or a pncu_status / cudaError. */

/* AtopBinaryMathFunction element-wise branch + AtopCompareMathFunction
 * (src/pnumpy/_pnumpy.cpp:732-791, 805-866) */
PNCU_API int nte_binary(pncu_any_two_fn launcher, int in_itemsize, int out_itemsize,
                        const void* in1, const void* in2, void* out, int64_t len,
                        int64_t stride1, int64_t stride2, int64_t strideOut, int max_workers);
/* AtopUnaryMathFunction + AtopTrigMathFunction (src/pnumpy/_pnumpy.cpp:874-996) */
PNCU_API int nte_unary(pncu_unary_fn launcher, int in_itemsize, int out_itemsize,
                       const void* in, void* out, int64_t len, int64_t strideIn, int64_t strideOut,
                       int max_workers);
/* reduce branch of AtopBinaryMathFunction (src/pnumpy/_pnumpy.cpp:646-731): *out = fold(in) op *out
 * when use_out_as_start, else fold(in) alone.  `out` is a HOST pointer to one element. */
PNCU_API int nte_reduce(int func, int atype, const void* in, void* out, int use_out_as_start,
                        int64_t len, int64_t strideIn, int max_workers);
/* nte_unary with the op's cost class (AtopTrigMathFunction passes NTE_CLASS_HEAVY) */
PNCU_API int nte_unary_class(pncu_unary_fn launcher, int in_itemsize, int out_itemsize,
                             const void* in, void* out, int64_t len, int64_t strideIn, int64_t strideOut,
                             int max_workers, int op_class);
/* AtopArgMinMaxMathFunction (src/pnumpy/_pnumpy.cpp:1105-1118); `in` contiguous */
PNCU_API int nte_argminmax(int kind, int atype, const void* in, int64_t len, int64_t* out_index,
                           int max_workers);
/* Fused chain: out = in1*in2 + in3 and *out = sum(in1*in2 + in3) over three
 * contiguous arrays of one type, bit-identical to the two / three hooked calls they replace
 * (pnumpy_cuda.h: pncu_any_three_fn, pncu_three_reduce_fn) with 16 / 12 bytes per float32 element
 * over PCIe or HBM instead of 24 / 28.  Same sharding, staging, residency and NTE_DECLINED rules as
 * the calls above; for nte_muladd_sum `out` is a HOST pointer to one element. */
PNCU_API int nte_muladd(int atype, const void* in1, const void* in2, const void* in3, void* out,
                        int64_t len, int max_workers);
PNCU_API int nte_muladd_sum(int atype, const void* in1, const void* in2, const void* in3, void* out,
                            int64_t len, int max_workers);

/* ---- pinned host memory for ndarrays (the recycler's allocator) ---------------------
 * Pinned + portable + mapped; operands that live in it are DMA'd in place, with no staging copy. */
PNCU_API void* nte_alloc_pinned(size_t bytes);
PNCU_API void nte_free_pinned(void* ptr);

/* ---- device-resident ndarrays: CUDA managed memory -------------------------------------------------
 * Host code (NumPy) and kernels share one pointer; pages migrate on demand.  The dispatcher runs
 * resident operands IN PLACE (no staging, no PCIe in steady state): it prefetches managed operands
 * to the GPU before the launch, so a ufunc chain over managed arrays touches PCIe only when the
 * host reads a result. */
PNCU_API void* nte_alloc_managed(size_t bytes);
PNCU_API void nte_free_managed(void* ptr);
/* zero a managed block on the GPU (its pages end up resident on GPU 0); 0 or a status */
PNCU_API int nte_zero_managed(void* ptr, size_t bytes);
/* forget where the pages of a managed block were last placed (the next call prefetches them again) */
PNCU_API void nte_forget_residency(void* ptr);

/* ---- accounting (feeds atop_info()/ledger and the tests) ------------------------------------------- */
typedef struct nte_stats {
    int64_t calls;          /* dispatches that ran on the GPU */
    int64_t declined;       /* dispatches handed back to the caller */
    int64_t launches;       /* kernels launched on behalf of those calls */
    int64_t h2d_bytes;      /* bytes copied host -> device */
    int64_t d2h_bytes;      /* bytes copied device -> host */
    int64_t staged_bytes;   /* bytes memcpy'd through pinned staging slabs (pageable operands) */
    int64_t lanes_last;     /* lanes used by the last call */
    int64_t gpus_last;      /* distinct GPUs used by the last call */
    int64_t prefetches;     /* cudaMemPrefetchAsync calls issued for managed operands */
} nte_stats;
PNCU_API void nte_get_stats(nte_stats* out);
/* per-call trace for latency work: when on, the dispatcher records CLOCK_MONOTONIC ns at eight stations of every
 * RESIDENT dispatch -- enter, lock taken, shards planned, operands placed, workers posted, own launch done (or own
 * shard finished, element-wise), workers joined, result in place -- and nte_get_trace copies the last call's eight
 * values (tools/resident_probe.py --trace prints the differences). */
PNCU_API void nte_set_trace(int on);
PNCU_API void nte_get_trace(int64_t* out8);
PNCU_API void nte_reset_stats(void);

#ifdef __cplusplus
}
#endif
#endif /* PNUMPY_NTE_H */
