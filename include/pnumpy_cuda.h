/*
 * pnumpy_cuda.h -- C ABI of libpnumpy_cuda, the sm_100a (B200) replacement for
 * pnumpy's `atop` loop-body library.
 *
 * What this replaces
 * ------------------
 * The reference's inner C ABI is `src/atop/atop.h:16-36`: `atop_init()` and a
 * family of getters (`GetSimpleMathOpFast`, `GetReduceMathOpFast`,
 * `GetComparisonOpFast`, `GetUnaryOpFast/Slow`, `GetTrigOpFast`) which hand out
 * AVX2 loop-body function pointers of the types declared in
 * `src/atop/common_inc.h:382-387`:
 *
 *     void UNARY  (void* in,  void* out, int64 len, int64 strideIn, int64 strideOut);
 *     void ANY_TWO(void* in1, void* in2, void* out, int64 len, int64 s1, int64 s2, int64 sOut);
 *     void REDUCE (void* in,  void* out, void* startVal, int64 len, int64 strideIn);
 *
 * Every entry point below mirrors one of those (same name with a `pncu_` prefix,
 * same enum VALUES, same argument meaning, strides in BYTES, stride 0 = scalar
 * broadcast) and differs in exactly three ways:
 *   1. the data pointers are DEVICE pointers (or any pointer the current device
 *      can dereference: cudaMalloc, cudaMallocManaged, mapped pinned host);
 *   2. each launcher takes a trailing `pncu_stream_t` and is ASYNCHRONOUS on it;
 *   3. each launcher returns an int status (0 = ok) instead of void, with the
 *      message available from `pncu_last_error()`.
 * There is no CPU fallback anywhere behind this header: with no usable CUDA
 * device every call fails with PNCU_ERR_NO_DEVICE.
 *
 * Plain C, no torch / numpy / Python types.  All symbols have C linkage.
 */
#ifndef PNUMPY_CUDA_H
#define PNUMPY_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define PNCU_API __attribute__((visibility("default")))
#else
#define PNCU_API
#endif

/* ---------------------------------------------------------------------------
 * Vocabulary.  The numeric VALUES are those of the reference enums and are baked
 * into its (and our) loop-stub tables; do not renumber.
 * ------------------------------------------------------------------------- */

/* reference: src/atop/common_inc.h:236-248 (ATOP_TYPES) */
enum pncu_atype {
    PNCU_BOOL = 0,
    PNCU_INT8 = 1, PNCU_UINT8 = 2,
    PNCU_INT16 = 3, PNCU_UINT16 = 4,
    PNCU_INT32 = 5, PNCU_UINT32 = 6,
    PNCU_INT64 = 7, PNCU_UINT64 = 8,
    PNCU_INT128 = 9, PNCU_UINT128 = 10,
    PNCU_HALF_FLOAT = 11, PNCU_FLOAT = 12, PNCU_DOUBLE = 13, PNCU_LONGDOUBLE = 14,
    PNCU_CHALF_FLOAT = 15, PNCU_CFLOAT = 16, PNCU_CDOUBLE = 17, PNCU_CLONGDOUBLE = 18,
    PNCU_STRING = 19, PNCU_UNICODE = 20,
    PNCU_VOID = 21,
    PNCU_ATYPE_LAST = 22
};

/* reference: src/atop/common_inc.h:250-259 (COMP_OPERATION) */
enum pncu_comp_op {
    PNCU_CMP_EQ = 0, PNCU_CMP_NE = 1, PNCU_CMP_LT = 2, PNCU_CMP_GT = 3,
    PNCU_CMP_LTE = 4, PNCU_CMP_GTE = 5, PNCU_CMP_LAST = 6
};

/* reference: src/atop/common_inc.h:261-303 (UNARY_OPERATION) */
enum pncu_unary_op {
    PNCU_UNARY_INVALID = 0,
    PNCU_ABS = 1, PNCU_SIGNBIT = 2, PNCU_FABS = 3, PNCU_INVERT = 4, PNCU_FLOOR = 5,
    PNCU_CEIL = 6, PNCU_TRUNC = 7, PNCU_ROUND = 8, PNCU_NEGATIVE = 9, PNCU_POSITIVE = 10,
    PNCU_SIGN = 11, PNCU_RINT = 12,
    PNCU_SQRT = 15, PNCU_SQUARE = 16, PNCU_RECIPROCAL = 17,
    PNCU_LOGICAL_NOT = 18, PNCU_ISINF = 19, PNCU_ISNAN = 20, PNCU_ISFINITE = 21,
    PNCU_ISNORMAL = 22, PNCU_ISNOTINF = 23, PNCU_ISNOTNAN = 24, PNCU_ISNOTFINITE = 25,
    PNCU_ISNOTNORMAL = 26, PNCU_ISNANORZERO = 27,
    PNCU_BITWISE_NOT = 28,
    PNCU_UNARY_LAST = 35
};

/* reference: src/atop/common_inc.h:305-346 (BINARY_OPERATION) */
enum pncu_binary_op {
    PNCU_BINARY_INVALID = 0,
    PNCU_ADD = 1, PNCU_SUB = 2, PNCU_MUL = 3, PNCU_MOD = 4,
    PNCU_MIN = 5, PNCU_MAX = 6, PNCU_NANMIN = 7, PNCU_NANMAX = 8,
    PNCU_FLOORDIV = 9, PNCU_POWER = 10, PNCU_REMAINDER = 11, PNCU_FMOD = 12,
    PNCU_DIV = 13, PNCU_SUBDATETIMES = 14, PNCU_SUBDATES = 15,
    PNCU_LOGICAL_AND = 16, PNCU_LOGICAL_XOR = 17, PNCU_LOGICAL_OR = 18,
    PNCU_BITWISE_LSHIFT = 19, PNCU_BITWISE_RSHIFT = 20, PNCU_BITWISE_AND = 21,
    PNCU_BITWISE_XOR = 22, PNCU_BITWISE_OR = 23, PNCU_BITWISE_ANDNOT = 24,
    PNCU_BITWISE_NOTAND = 25, PNCU_BITWISE_XOR_SPECIAL = 26,
    PNCU_ATAN2 = 27, PNCU_HYPOT = 28,
    PNCU_BINARY_LAST = 29
};

/* reference: src/atop/common_inc.h:348-372 (TRIG_OPERATION) */
enum pncu_trig_op {
    PNCU_TRIG_INVALID = 0,
    PNCU_SIN = 1, PNCU_COS = 2, PNCU_TAN = 3, PNCU_ASIN = 4, PNCU_ACOS = 5, PNCU_ATAN = 6,
    PNCU_SINH = 7, PNCU_COSH = 8, PNCU_TANH = 9, PNCU_ASINH = 10, PNCU_ACOSH = 11, PNCU_ATANH = 12,
    PNCU_LOG = 13, PNCU_LOG2 = 14, PNCU_LOG10 = 15, PNCU_EXP = 16, PNCU_EXP2 = 17,
    PNCU_EXPM1 = 18, PNCU_LOG1P = 19, PNCU_CBRT = 20,
    PNCU_TRIG_LAST = 21
};

/* arg-reduction kind; values are the reference's gArgMinMaxMapping
 * (src/pnumpy/_pnumpy.cpp:186-189): argmax = 0, argmin = 1 */
enum pncu_argminmax_kind { PNCU_ARGMAX = 0, PNCU_ARGMIN = 1 };

/* status codes: 0 ok, >0 a cudaError_t value, <0 ours */
enum pncu_status {
    PNCU_OK = 0,
    PNCU_ERR_NO_DEVICE   = -1,  /* no sm_100 CUDA device / driver */
    PNCU_ERR_UNSUPPORTED = -2,  /* op/type/layout has no kernel */
    PNCU_ERR_OVERLAP     = -3,  /* out partially overlaps an input (not exact aliasing) */
    PNCU_ERR_ARGUMENT    = -4,
    PNCU_ERR_NOMEM       = -5,
    PNCU_ERR_TIMEOUT     = -6   /* a participant of a cross-GPU exchange never arrived; no result was produced */
};

/* opaque: a cudaStream_t (NULL = the legacy default stream of the current device) */
typedef void* pncu_stream_t;
typedef void* pncu_event_t;

/* ---------------------------------------------------------------------------
 * Launcher types == the reference's loop-body pointer types + stream + status.
 * reference: src/atop/common_inc.h:382-387
 * ------------------------------------------------------------------------- */

/* UNARY_FUNC: out[i] = f(in[i]) */
typedef int (*pncu_unary_fn)(const void* in, void* out, int64_t len,
                             int64_t strideIn, int64_t strideOut, pncu_stream_t stream);

/* ANY_TWO_FUNC: out[i] = in1[i] op in2[i]; a stride of 0 broadcasts that operand */
typedef int (*pncu_any_two_fn)(const void* in1, const void* in2, void* out, int64_t len,
                               int64_t strideIn1, int64_t strideIn2, int64_t strideOut,
                               pncu_stream_t stream);

/* REDUCE_FUNC: *out = fold(op, in[0..len)) op *startVal.
 * `startVal` may alias `out` (NumPy seeds the accumulator that way,
 * reference: src/atop/ops_binary.cpp:199-200) or be NULL (= the op's identity;
 * for MIN/MAX the first element).  `out` and `startVal` are device pointers to ONE
 * element of the op's type.  Deterministic: the result depends on (in, len) only,
 * never on grid size, SM count or scheduling. */
typedef int (*pncu_reduce_fn)(const void* in, void* out, const void* startVal, int64_t len,
                              int64_t strideIn, pncu_stream_t stream);

/* arg-reduction behind NumPy's PyArray_ArgFunc slot
 * (reference: src/pnumpy/_pnumpy.cpp:1105-1118, stubs.h:619-676).
 * `in` contiguous, `out_index` a device pointer to one int64.  First occurrence
 * wins; for floats the first NaN wins (NumPy semantics). */
typedef int (*pncu_argminmax_fn)(const void* in, int64_t len, int64_t* out_index,
                                 pncu_stream_t stream);

/* Fused chain (the reference runs `a*b + c` as two ufunc calls through
 * SimpleMathOpFast, src/atop/ops_binary.cpp:398-571, and `.sum()` as a third through
 * ReduceMathOpFast, :186-254).  Three contiguous inputs of one type.
 * pncu_any_three_fn:    out[i] = (in1[i] * in2[i]) + in3[i]; the product and the sum are rounded
 *                       separately (no FMA), so the result is bit-identical to multiply-then-add.
 * pncu_three_reduce_fn: *out = sum_i ((in1[i] * in2[i]) + in3[i]) with the block structure, fold
 *                       order and accumulator type of the ADD reduction, so it is bit-identical to
 *                       multiply, add, then sum (operands 32-byte aligned alike) while HBM is
 *                       read once.  `out` is a device pointer to one element. */
typedef int (*pncu_any_three_fn)(const void* in1, const void* in2, const void* in3, void* out,
                                 int64_t len, pncu_stream_t stream);
typedef int (*pncu_three_reduce_fn)(const void* in1, const void* in2, const void* in3, void* out,
                                    int64_t len, pncu_stream_t stream);

/* ---------------------------------------------------------------------------
 * Library lifetime.  replaces: atop_init() (src/atop/atop.cpp:68-104)
 * ------------------------------------------------------------------------- */

/* Probe the driver and enumerate devices.  Returns 0 and writes the device count,
 * or PNCU_ERR_NO_DEVICE.  Idempotent and thread-safe. */
PNCU_API int pncu_init(int* device_count);
PNCU_API int pncu_shutdown(void);
PNCU_API int pncu_device_count(void);
/* "NVIDIA B200 sm_100 148 SMs 178.4 GiB ..." -- feeds cpustring()/atop_info
 * (reference: CMathWorker::CPUString, src/atop/threads.cpp:616-656) */
PNCU_API int pncu_device_info(int device, char* buf, size_t buflen);
/* message of the last failing call on this host thread ("" if none) */
PNCU_API const char* pncu_last_error(void);
/* number of kernels this library has launched since load (all threads) */
PNCU_API int64_t pncu_launch_count(void);
/* run-time tunables ("ctas_per_sm", "block", "reduce_ctas_per_sm"); returns previous or <0 */
PNCU_API int pncu_set_tunable(const char* name, int value);

/* ---------------------------------------------------------------------------
 * Getters.  Same contract as the reference's: return NULL when there is no kernel
 * for (func, type); write *wantedOutType only when the (func, type) pair is one the
 * hook layer must register (src/pnumpy/_pnumpy.cpp:1349,1464) and otherwise leave
 * it untouched.
 * ------------------------------------------------------------------------- */

/* replaces GetSimpleMathOpFast  (src/atop/ops_binary.cpp:871-1058) */
PNCU_API pncu_any_two_fn pncu_GetSimpleMathOpFast(int func, int atopInType1, int atopInType2,
                                                  int* wantedOutType);
/* replaces GetReduceMathOpFast  (src/atop/ops_binary.cpp:1060-1127) */
PNCU_API pncu_reduce_fn pncu_GetReduceMathOpFast(int func, int atopInType1);
/* replaces GetComparisonOpFast  (src/atop/ops_compare.cpp:1266-1446) */
PNCU_API pncu_any_two_fn pncu_GetComparisonOpFast(int func, int atopInType1, int atopInType2,
                                                  int* wantedOutType);
/* replaces GetUnaryOpFast + GetUnaryOpSlow (src/atop/ops_unary.cpp:1250-1774);
 * the out-type written is what the reference's Fast-then-Slow sequence leaves behind */
PNCU_API pncu_unary_fn pncu_GetUnaryOpFast(int func, int atopInType1, int* wantedOutType);
/* replaces GetTrigOpFast        (src/atop/ops_trig.cpp:578-781) */
PNCU_API pncu_unary_fn pncu_GetTrigOpFast(int func, int atopInType1, int* wantedOutType);
/* new: the reference forwards arg-reductions to NumPy (src/pnumpy/_pnumpy.cpp:1105-1118) */
PNCU_API pncu_argminmax_fn pncu_GetArgMinMaxOpFast(int kind, int atopInType1);
/* int/int true_divide producing float64 with cvt.rn + div.rn.  The reference does not
 * hook it (src/atop/ops_binary.cpp:926); NumPy casts to the hooked dd->d loop.  Offered
 * for device-resident callers so the cast never touches HBM twice. */
PNCU_API pncu_any_two_fn pncu_GetIntTrueDivide(int atopInType1);
/* fused a*b+c and sum(a*b+c) for int32/uint32/int64/uint64/float32/float64; NULL otherwise */
PNCU_API pncu_any_three_fn pncu_GetMulAddFast(int atopInType1);
PNCU_API pncu_three_reduce_fn pncu_GetMulAddSumFast(int atopInType1);
/* dtype conversion (`astype`), the first item of the reference's roadmap (doc_src/source/roadmap.rst:11-15;
 * SURVEY.md 8f row f-4): out[i] = (OutType)in[i] with NumPy's cast semantics -- int -> int wraps, anything -> bool
 * is x != 0, int -> float and float64 -> float32 round to nearest even, float -> int truncates toward zero and
 * returns what x86's cvttss2si / cvttsd2si return when the value does not fit (NumPy does the same and warns).
 * A pncu_unary_fn over (inType, outType) for bool, the eight integer types, float32 and float64; NULL otherwise. */
PNCU_API pncu_unary_fn pncu_GetConvertFast(int atopInType, int atopOutType);
/* itemsize of an atop type, 0 if unsupported (src/pnumpy/_pnumpy.cpp:43-54) */
PNCU_API int pncu_itemsize(int atype);

/* ---------------------------------------------------------------------------
 * Reduction building blocks for the multi-GPU dispatcher (pnumpy_nte.h).
 *
 * A reduction over [0, n) of a type with `itemsize` bytes is defined on fixed LOGICAL BLOCKS of
 * 64 KiB, i.e. 65536 / itemsize elements: block k covers [k*B, min((k+1)*B, n)) and is folded in
 * a fixed tree into partial slot k; the block partials are then folded in index order by ONE
 * thread block.  Because B depends on the type only -- not on the device, the grid or the number
 * of shards -- a shard boundary that is a multiple of pncu_reduce_block_elems() (= 65536, a
 * multiple of every type's B) gives bit-identical results on 1, 2, 4 or 8 GPUs.  This is the GPU
 * form of the reference's per-16K-chunk partials + second pass on the caller
 * (src/pnumpy/_pnumpy.cpp:666-700).
 * ------------------------------------------------------------------------- */
/* alignment (in elements) that shard boundaries must have */
PNCU_API int64_t pncu_reduce_block_elems(void);
/* pass 2 folds the slots (block partials) in groups of this many consecutive slots first, then the group partials */
PNCU_API int pncu_reduce_group_slots(void);
/* number of partial slots pass 1 writes for `len` elements of `atype` (= ceil(len / B)) */
PNCU_API int64_t pncu_reduce_partial_count(int atype, int64_t len);
/* bytes of one block partial for (func, type); 0 if no kernel */
PNCU_API int pncu_reduce_partial_bytes(int func, int atype);
/* pass 1: partials[k] for k in [0, pncu_reduce_partial_count(atype, len)) ; `in` contiguous */
PNCU_API int pncu_reduce_partials(int func, int atype, const void* in, int64_t len,
                                  void* partials, pncu_stream_t stream);
/* pass 2: *out = fold(partials[0..count)) op *startVal (startVal may be NULL / alias out) */
PNCU_API int pncu_reduce_finish(int func, int atype, const void* partials, int64_t count,
                                void* out, const void* startVal, pncu_stream_t stream);
/* pass 1 of sum(in1*in2 + in3); the partials have the layout of pncu_reduce_partials(PNCU_ADD,
 * atype, ...) and are finished by pncu_reduce_finish(PNCU_ADD, atype, ...) */
PNCU_API int pncu_muladd_sum_partials(int atype, const void* in1, const void* in2, const void* in3,
                                      int64_t len, void* partials, pncu_stream_t stream);
/* arg-reduction equivalents; a partial is {value bits (8 B), index (8 B)}.
 * `index_base` is added to every index (the shard's global offset). */
PNCU_API int pncu_argminmax_partials(int kind, int atype, const void* in, int64_t len,
                                     int64_t index_base, void* partials, pncu_stream_t stream);
PNCU_API int pncu_argminmax_finish(int kind, int atype, const void* partials, int64_t count,
                                   int64_t* out_index, pncu_stream_t stream);

/* ---------------------------------------------------------------------------
 * The cross-GPU exchange fused into pass 2 (the path's ONE exchange step).
 *
 * The reference folds its per-chunk partials on the calling thread (src/pnumpy/_pnumpy.cpp:697-699); with
 * the input sharded over G GPUs the block partials of all shards have to meet.  Instead of a collective
 * between the passes, each participant's pass 2 is ONE kernel that (i) stores its own block partials into
 * slot `slot_base ..` of EVERY participant's slot table -- ordinary stores into peer memory over NVLink,
 * pncu_exchange_slices() CTAs per destination, each followed by a system-scope fence and one arrival count
 * on the destination's counter -- (ii) waits ON THE DEVICE until the slices of all participants have
 * arrived in its own table and (iii) folds the complete table in slot order.  Every GPU ends with the same
 * bits (an all-reduce), identical to the single-GPU result, with no NCCL call, no extra launch and no host
 * synchronisation between the passes.
 *
 * Tables: `slots[d]` points to participant d's slot array (element size pncu_reduce_partial_bytes(), or 16
 * for arg-reductions), `arrived[d]` to its 64-bit counter; both device memory of GPU d that the calling
 * GPU can write (same process: pncu_enable_peer_access; other processes: pncu_ipc_export / pncu_ipc_import).
 * Counters only ever grow: every call adds n * pncu_exchange_slices() arrivals to each table, and the
 * caller passes the running total as `expected`.  Alternate between two tables on successive calls (a fast
 * rank may start call e+1 while a slow one still folds the table of call e).
 * ------------------------------------------------------------------------- */
#define PNCU_MAX_PEERS 8
typedef struct pncu_peer_table {
    int n;                                        /* participants, 1..PNCU_MAX_PEERS */
    void* slots[PNCU_MAX_PEERS];
    unsigned long long* arrived[PNCU_MAX_PEERS];
} pncu_peer_table;
PNCU_API int pncu_exchange_slices(void);
/* `local_partials`: the output of pncu_reduce_partials() on this participant's shard (local_count slots);
 * `self`: this participant's index in `peers`; `total_count`: slots of all participants together.
 * `status` (optional, device-visible memory, zero-initialised): when a participant has not arrived within the
 * bounded wait (~2 s) the kernel does NOT fold the incomplete table and does NOT write `out`; it ORs
 * PNCU_FUSED_TIMEOUT_BIT into *status and counts the event (pncu_exchange_timeouts).  The caller must check
 * one of the two before using `out`. */
PNCU_API int pncu_reduce_exchange_finish(int func, int atype, const void* local_partials, int64_t local_count,
                                         int64_t slot_base, const pncu_peer_table* peers, int self,
                                         int64_t total_count, void* out, const void* startVal,
                                         unsigned long long expected, unsigned long long* status,
                                         pncu_stream_t stream);
PNCU_API int pncu_argminmax_exchange_finish(int kind, int atype, const void* local_partials, int64_t local_count,
                                            int64_t slot_base, const pncu_peer_table* peers, int self,
                                            int64_t total_count, int64_t* out_index,
                                            unsigned long long expected, unsigned long long* status,
                                            pncu_stream_t stream);
/* bounded waits that gave up (a participant never arrived) on the current device; 0 in a healthy job */
PNCU_API int64_t pncu_exchange_timeouts(void);

/* ---------------------------------------------------------------------------
 * ONE-LAUNCH reductions: pass 1, the exchange and pass 2 in a single kernel per GPU.
 *
 * Every pass-1 CTA takes a ticket after storing its block partial; the CTA that draws the last ticket does
 * pass 2 itself, in the fold order of pncu_reduce_finish, so the result has the same bits as the two-launch
 * building blocks above (this is also what the pncu_GetReduceMathOpFast / pncu_GetArgMinMaxOpFast /
 * pncu_GetMulAddSumFast launchers run).  With `x->n > 1` participants (one shard per GPU, shard boundaries at
 * multiples of pncu_reduce_block_elems()), the last CTA of every participant first stores the shard's partials
 * into slot `slot_base ..` of each CONSUMER's table over NVLink (one system-scope fence, one arrival count per
 * consumer) and a consumer's last CTA waits on the device for the other n-1 shards before it folds the complete
 * table of `total_count` slots.  Consumers: every participant when `all` (an all-reduce: what one process per
 * GPU wants), else only `root` (what a single process driving n GPUs wants).  A consumer's own pass 1 writes
 * straight into its table, so `slots[self]` must be memory of the launching GPU.  `arrived[d]` only ever grows:
 * each call adds n-1 to every consumer's counter and passes the running total as `expected`.  When calls can
 * overlap in time (all-reduce across processes) alternate between two tables, as for pncu_reduce_exchange_finish.
 * `out` / `startVal` may be device memory or MAPPED pinned host memory; `done_flag` (optional, device or mapped
 * host memory) receives `done_value` with a system-scope release store after `out` is written, so a host thread
 * can poll it instead of synchronising the stream.  If a shard never arrives (~2 s) the consumer does not fold
 * and does not write `out`; it publishes `done_value | PNCU_FUSED_TIMEOUT_BIT` and counts the event
 * (pncu_exchange_timeouts).  After such a failure call pncu_fused_reset(stream) before reusing the stream.
 * Pass 2 folds the slots in two levels (groups of pncu_reduce_group_slots() = 64 consecutive slots, then the group
 * partials).  When EVERY participant's `slot_base` is a multiple of 64 and `group_offset` is set (every consumer's
 * allocation `slots[d]` then holds a second table of (total_count + 63) / 64 partials at that byte offset, behind
 * the slots), each participant folds its own groups and stores only those: 1/64 of the exchange traffic and no
 * level-A work left for the consumer.  Same bits either way.
 * x == NULL: a single GPU with no completion flag.
 * replaces: the per-chunk partials + second pass on the caller, src/pnumpy/_pnumpy.cpp:666-700.
 * ------------------------------------------------------------------------- */
#define PNCU_FUSED_TIMEOUT_BIT (1ULL << 63)
typedef struct pncu_fused_exchange {
    int n;                                        /* participants, 1..PNCU_MAX_PEERS (1 = no exchange) */
    int self;                                     /* this participant */
    int all;                                      /* 1: every participant folds; 0: only `root` */
    int root;
    int64_t slot_base;                            /* first slot of this participant's shard */
    int64_t total_count;                          /* slots of all shards together */
    void* slots[PNCU_MAX_PEERS];                  /* consumers' tables (others may be NULL) */
    unsigned long long* arrived[PNCU_MAX_PEERS];  /* consumers' arrival counters */
    unsigned long long expected;                  /* running total a consumer's counter reaches with this call */
    unsigned long long* done_flag;                /* optional */
    unsigned long long done_value;
    int64_t group_offset;                         /* optional (0 = none): byte offset of the GROUP table in every consumer's `slots[d]` */
} pncu_fused_exchange;
PNCU_API int pncu_reduce_fused(int func, int atype, const void* in, int64_t len, const pncu_fused_exchange* x,
                               void* out, const void* startVal, pncu_stream_t stream);
PNCU_API int pncu_argminmax_fused(int kind, int atype, const void* in, int64_t len, int64_t index_base,
                                  const pncu_fused_exchange* x, int64_t* out_index, pncu_stream_t stream);
PNCU_API int pncu_muladd_sum_fused(int atype, const void* in1, const void* in2, const void* in3, int64_t len,
                                   const pncu_fused_exchange* x, void* out, pncu_stream_t stream);
/* profiling aid: nanoseconds the most recent one-launch reduction on the current device spent in its tail (the last
 * CTA: ticket drawn -> exchange -> pass 2 -> result published); synchronises the device */
PNCU_API int64_t pncu_last_tail_ns(void);
/* re-arm the stream's ticket counter after a failed or timed-out one-launch reduction */
PNCU_API int pncu_fused_reset(pncu_stream_t stream);
/* peer memory plumbing: same-process peer access, and CUDA IPC handles (64 bytes) for one-process-per-GPU jobs */
PNCU_API int pncu_enable_peer_access(int peer_device);     /* from the current device; idempotent */
PNCU_API int pncu_ipc_export(void* dev_ptr, void* handle64);
PNCU_API int pncu_ipc_import(const void* handle64, void** dev_ptr);
PNCU_API int pncu_ipc_close(void* dev_ptr);

/* ---------------------------------------------------------------------------
 * Device memory / stream / event helpers (thin cudart wrappers) so that callers
 * written in C, ctypes or cgo need no CUDA headers.
 * ------------------------------------------------------------------------- */
PNCU_API int pncu_set_device(int device);
PNCU_API int pncu_get_device(int* device);
PNCU_API int pncu_malloc(void** ptr, size_t bytes);           /* on the current device */
PNCU_API int pncu_free(void* ptr);
PNCU_API int pncu_malloc_host(void** ptr, size_t bytes);      /* pinned, portable, mapped */
PNCU_API int pncu_free_host(void* ptr);
PNCU_API int pncu_host_register(void* ptr, size_t bytes);     /* pin an existing host range */
PNCU_API int pncu_host_unregister(void* ptr);
PNCU_API int pncu_memcpy_h2d(void* dst, const void* src, size_t bytes, pncu_stream_t stream);
PNCU_API int pncu_memcpy_d2h(void* dst, const void* src, size_t bytes, pncu_stream_t stream);
PNCU_API int pncu_memcpy_d2d(void* dst, const void* src, size_t bytes, pncu_stream_t stream);
PNCU_API int pncu_memset(void* dst, int value, size_t bytes, pncu_stream_t stream);
PNCU_API int pncu_stream_create(pncu_stream_t* stream);       /* non-blocking stream */
PNCU_API int pncu_stream_destroy(pncu_stream_t stream);
PNCU_API int pncu_stream_sync(pncu_stream_t stream);
PNCU_API int pncu_device_sync(void);
PNCU_API int pncu_event_create(pncu_event_t* ev);
PNCU_API int pncu_event_destroy(pncu_event_t ev);
PNCU_API int pncu_event_record(pncu_event_t ev, pncu_stream_t stream);
PNCU_API int pncu_event_sync(pncu_event_t ev);
PNCU_API int pncu_event_elapsed_ms(pncu_event_t start, pncu_event_t stop, float* ms);
/* 0 = host pageable/unknown, 1 = pinned host, 2 = device, 3 = managed; device id in *device */
PNCU_API int pncu_pointer_kind(const void* ptr, int* device);

#ifdef __cplusplus
}
#endif
#endif /* PNUMPY_CUDA_H */
