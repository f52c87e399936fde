// _pnumpy -- the CPython extension that installs the B200 loops into NumPy.
//
// replaces: src/pnumpy/_pnumpy.cpp (hook registration `newinit` :1287-1676, the four per-call
// dispatchers :638-996, the arg-reduction dispatcher :1105-1118, the control functions
// :1766-1966) and src/pnumpy/module_init.cpp:60-124 (method table, module init).
//
// Same hooks, same tables, same Python surface.  What changed underneath:
//   * the loop bodies come from libpnumpy_cuda's getters (pncu_Get*OpFast) instead of atop's,
//   * a call is handed to the nte dispatcher (nte_binary / nte_unary / nte_reduce / nte_argminmax)
//     instead of THREADER->WorkMain; nte stages host operands through pinned slabs, shards them over
//     the GPUs and returns when the result is in the caller's buffers,
//   * when nte declines (small n, odd strides, re-entrant call) or atop is disabled, the saved
//     NumPy loop `pOldFunc` runs -- exactly where the reference runs it
//     (src/pnumpy/_pnumpy.cpp:653-663, 748-759).
// A CUDA failure is never papered over: it raises RuntimeError.
#define NPY_NO_DEPRECATED_API NPY_1_7_API_VERSION
#define NPY_TARGET_VERSION NPY_2_0_API_VERSION  /* PyUFunc_AddLoopFromSpec, PyDataType_GetArrFuncs */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <numpy/arrayobject.h>
#include <numpy/ufuncobject.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <map>
#include <mutex>
#include <vector>

#include "pnumpy_nte.h"

#include "stubs.h"

// ---- dtype <-> atop tables (reference: src/pnumpy/_pnumpy.cpp:7-54; Linux LP64 column) ------------------
static const int convert_dtype_to_atop[] = {
    PNCU_BOOL,                              // NPY_BOOL = 0
    PNCU_INT8, PNCU_UINT8,                  // NPY_BYTE, NPY_UBYTE
    PNCU_INT16, PNCU_UINT16,                // NPY_SHORT, NPY_USHORT
    PNCU_INT32, PNCU_UINT32,                // NPY_INT, NPY_UINT
    PNCU_INT64, PNCU_UINT64,                // NPY_LONG, NPY_ULONG
    PNCU_INT64, PNCU_UINT64,                // NPY_LONGLONG, NPY_ULONGLONG
    PNCU_FLOAT, PNCU_DOUBLE, PNCU_LONGDOUBLE,
    PNCU_CFLOAT, PNCU_CDOUBLE, PNCU_CLONGDOUBLE,
    -1,                                     // NPY_OBJECT
    PNCU_STRING, PNCU_UNICODE, PNCU_VOID,
    -1, -1, PNCU_HALF_FLOAT                 // NPY_DATETIME, NPY_TIMEDELTA, NPY_HALF
};
static const int convert_atop_to_dtype[] = {
    NPY_BOOL, NPY_INT8, NPY_UINT8, NPY_INT16, NPY_UINT16, NPY_INT32, NPY_UINT32, NPY_INT64, NPY_UINT64,
    NPY_LONGLONG, NPY_ULONGLONG,  // "really INT128" in the reference
    NPY_HALF, NPY_FLOAT, NPY_DOUBLE, NPY_LONGDOUBLE,
    NPY_HALF, NPY_CFLOAT, NPY_CDOUBLE, NPY_CLONGDOUBLE,
    NPY_STRING, NPY_UNICODE, NPY_VOID
};

struct stUFuncToAtop { const char* str_ufunc_name; int atop_op; };

// reference: gBinaryMapping / gCompareMapping / gUnaryMapping / gTrigMapping, src/pnumpy/_pnumpy.cpp:90-170
static const stUFuncToAtop gBinaryMapping[] = {
    {"add", PNCU_ADD}, {"subtract", PNCU_SUB}, {"multiply", PNCU_MUL}, {"true_divide", PNCU_DIV},
    {"floor_divide", PNCU_FLOORDIV}, {"minimum", PNCU_MIN}, {"maximum", PNCU_MAX}, {"power", PNCU_POWER},
    {"remainder", PNCU_REMAINDER}, {"logical_and", PNCU_LOGICAL_AND}, {"logical_or", PNCU_LOGICAL_OR},
    {"bitwise_and", PNCU_BITWISE_AND}, {"bitwise_or", PNCU_BITWISE_OR}, {"bitwise_xor", PNCU_BITWISE_XOR},
    {"left_shift", PNCU_BITWISE_LSHIFT}, {"right_shift", PNCU_BITWISE_RSHIFT},
};
static const stUFuncToAtop gCompareMapping[] = {
    {"equal", PNCU_CMP_EQ}, {"not_equal", PNCU_CMP_NE}, {"greater", PNCU_CMP_GT},
    {"greater_equal", PNCU_CMP_GTE}, {"less", PNCU_CMP_LT}, {"less_equal", PNCU_CMP_LTE},
};
static const stUFuncToAtop gUnaryMapping[] = {
    {"abs", PNCU_ABS}, {"signbit", PNCU_SIGNBIT}, {"fabs", PNCU_FABS}, {"invert", PNCU_INVERT},
    {"floor", PNCU_FLOOR}, {"ceil", PNCU_CEIL}, {"trunc", PNCU_TRUNC}, {"rint", PNCU_ROUND},
    {"negative", PNCU_NEGATIVE}, {"positive", PNCU_POSITIVE}, {"sign", PNCU_SIGN}, {"rint", PNCU_RINT},
    {"sqrt", PNCU_SQRT}, {"square", PNCU_SQUARE}, {"reciprocal", PNCU_RECIPROCAL},
    {"logical_not", PNCU_LOGICAL_NOT}, {"isinf", PNCU_ISINF}, {"isnan", PNCU_ISNAN},
    {"isfinite", PNCU_ISFINITE}, {"bitwise_not", PNCU_BITWISE_NOT},
};
static const stUFuncToAtop gTrigMapping[] = {
    {"sin", PNCU_SIN}, {"cos", PNCU_COS}, {"tan", PNCU_TAN}, {"arcsin", PNCU_ASIN}, {"arccos", PNCU_ACOS},
    {"arctan", PNCU_ATAN}, {"sinh", PNCU_SINH}, {"cosh", PNCU_COSH}, {"tanh", PNCU_TANH},
    {"arcsinh", PNCU_ASINH}, {"arccosh", PNCU_ACOSH}, {"arctanh", PNCU_ATANH}, {"cbrt", PNCU_CBRT},
    {"exp", PNCU_EXP}, {"exp2", PNCU_EXP2}, {"expm1", PNCU_EXPM1}, {"log", PNCU_LOG}, {"log2", PNCU_LOG2},
    {"log10", PNCU_LOG10}, {"log1p", PNCU_LOG1P},
};
#define COUNT_OF(a) ((int)(sizeof(a) / sizeof((a)[0])))

// reference: stUFunc, src/pnumpy/_pnumpy.cpp:239-253 (same fields; pReduceFunc holds the launcher the
// GPU library would use for a device-resident caller, the hook itself goes through nte_reduce)
struct stUFunc {
    union {
        pncu_any_two_fn pBinaryFunc;
        pncu_unary_fn pUnaryFunc;
    };
    PyUFuncGenericFunction pOldFunc;
    pncu_reduce_fn pReduceFunc;
    int32_t MaxThreads;
    int32_t MinElementsToThread;
    int32_t OutType;  // atop type of the output
};
struct stArgMinMaxFunc {
    ArgMinMaxFunc pOldFunc;
    pncu_argminmax_fn pFunc;
    int32_t MaxThreads;
    int32_t MinElementsToThread;
};

static stUFunc g_UFuncLUT[PNCU_BINARY_LAST][PNCU_ATYPE_LAST];
static stUFunc g_CompFuncLUT[PNCU_CMP_LAST][PNCU_ATYPE_LAST];
static stUFunc g_UnaryFuncLUT[PNCU_UNARY_LAST][PNCU_ATYPE_LAST];
static stUFunc g_TrigFuncLUT[PNCU_TRIG_LAST][PNCU_ATYPE_LAST];
static stArgMinMaxFunc g_ArgMinMaxFuncLUT[2][PNCU_ATYPE_LAST];

// reference: stSettings, src/pnumpy/common.h:23-31, initial values src/pnumpy/_pnumpy.cpp:348
struct stSettings {
    int32_t AtopEnabled;
    int32_t LedgerEnabled;
    int32_t RecyclerEnabled;
    int32_t ZigZag;
    int32_t Initialized;
};
static stSettings g_Settings = {1, 0, 0, 0, 0};
static PyObject* gpUnaryDict = NULL;  // {(ufunc_name, dtype_num): address of the stUFunc}
static char g_DeviceString[512] = "";

// ---- ledger: per-category counters and a ring of per-call records while enabled (reference:
// src/pnumpy/ledger.cpp:144-272 keeps a ring of records -- name, lengths, start/end ticks -- and printf()s each
// one; the ring and the toggles are kept, the printf is not).  Ticks are timer_gettsc() values. -----------------
struct LedgerCounters { int64_t gpu_calls, old_calls, elements; };
static LedgerCounters g_Ledger[5];
enum { CAT_BINARY, CAT_UNARY, CAT_COMPARE, CAT_TRIG, CAT_ARGMINMAX };
struct LedgerRecord {
    uint64_t t0, t1;      // rdtsc at dispatcher entry / exit
    int64_t n;            // elements
    int64_t bytes;        // algorithmic bytes moved (inputs + output)
    int16_t cat, funcop, atype;
    int8_t gpu, reduce;
};
static const int kLedgerRing = 4096;
static LedgerRecord g_LedgerRing[kLedgerRing];
static uint64_t g_LedgerHead = 0;     // records ever written
static std::mutex g_LedgerMu;
static inline uint64_t rdtsc_now(void);
static inline void ledger(int cat, bool gpu, int64_t n, uint64_t t0 = 0, int funcop = 0, int atype = 0, int64_t bytes = 0,
                          bool reduce = false) {
    if (!g_Settings.LedgerEnabled) return;
    std::lock_guard<std::mutex> lk(g_LedgerMu);
    if (gpu) g_Ledger[cat].gpu_calls++; else g_Ledger[cat].old_calls++;
    g_Ledger[cat].elements += n;
    LedgerRecord& r = g_LedgerRing[g_LedgerHead++ % kLedgerRing];
    r.t0 = t0; r.t1 = rdtsc_now(); r.n = n; r.bytes = bytes;
    r.cat = (int16_t)cat; r.funcop = (int16_t)funcop; r.atype = (int16_t)atype; r.gpu = gpu; r.reduce = reduce;
}
static inline uint64_t ledger_t0(void) { return g_Settings.LedgerEnabled ? rdtsc_now() : 0; }

// A CUDA failure inside a loop: NumPy has dropped the GIL around large loops, so take it back first.
static void raise_cuda_error(const char* where, int rc) {
    PyGILState_STATE st = PyGILState_Ensure();
    if (!PyErr_Occurred())
        PyErr_Format(PyExc_RuntimeError, "pnumpy_b200: %s failed on the GPU (status %d): %s", where, rc, pncu_last_error());
    PyGILState_Release(st);
}

#define IS_BINARY_REDUCE ((args[0] == args[2]) && (steps[0] == steps[2]) && (steps[0] == 0))

// =======================================================================================================
// dispatchers
// reference: AtopBinaryMathFunction, src/pnumpy/_pnumpy.cpp:638-799
static void AtopBinaryMathFunction(char** args, const npy_intp* dimensions, const npy_intp* steps, void* innerloop, int funcop, int atype) {
    stUFunc* p = &g_UFuncLUT[funcop][atype];
    const npy_intp n = dimensions[0];
    const uint64_t t0 = ledger_t0();
    const bool reduce = IS_BINARY_REDUCE;
    const int isz = pncu_itemsize(atype);
    const int64_t bytes = reduce ? (int64_t)n * isz : (int64_t)n * (2 * isz + pncu_itemsize(p->OutType));
    if (g_Settings.AtopEnabled) {
        int rc = NTE_DECLINED;
        if (reduce) {
            // the middle array is the real input; the accumulator seeds the fold (:646-652)
            if (p->pReduceFunc) rc = nte_reduce(funcop, atype, args[1], args[0], 1, (int64_t)n, (int64_t)steps[1], p->MaxThreads);
        } else if (p->pBinaryFunc) {
            rc = nte_binary(p->pBinaryFunc, isz, pncu_itemsize(p->OutType), args[0], args[1], args[2],
                            (int64_t)n, (int64_t)steps[0], (int64_t)steps[1], (int64_t)steps[2], p->MaxThreads);
        }
        if (rc == 0) { ledger(CAT_BINARY, true, n, t0, funcop, atype, bytes, reduce); return; }
        if (rc != NTE_DECLINED) { raise_cuda_error("binary ufunc loop", rc); return; }
    }
    p->pOldFunc(args, dimensions, steps, innerloop);
    ledger(CAT_BINARY, false, n, t0, funcop, atype, bytes, reduce);
}

// reference: AtopCompareMathFunction, src/pnumpy/_pnumpy.cpp:805-866
static void AtopCompareMathFunction(char** args, const npy_intp* dimensions, const npy_intp* steps, void* innerloop, int funcop, int atype) {
    stUFunc* p = &g_CompFuncLUT[funcop][atype];
    const npy_intp n = dimensions[0];
    const uint64_t t0 = ledger_t0();
    const int64_t bytes = (int64_t)n * (2 * pncu_itemsize(atype) + 1);
    if (g_Settings.AtopEnabled && p->pBinaryFunc) {
        int rc = nte_binary(p->pBinaryFunc, pncu_itemsize(atype), 1, args[0], args[1], args[2], (int64_t)n,
                            (int64_t)steps[0], (int64_t)steps[1], (int64_t)steps[2], p->MaxThreads);
        if (rc == 0) { ledger(CAT_COMPARE, true, n, t0, funcop, atype, bytes); return; }
        if (rc != NTE_DECLINED) { raise_cuda_error("comparison ufunc loop", rc); return; }
    }
    p->pOldFunc(args, dimensions, steps, innerloop);
    ledger(CAT_COMPARE, false, n, t0, funcop, atype, bytes);
}

static void unary_common(stUFunc* p, int cat, char** args, const npy_intp* dimensions, const npy_intp* steps, void* innerloop, int atype,
                         int funcop, bool const_output = false) {
    const npy_intp n = dimensions[0];
    const uint64_t t0 = ledger_t0();
    const int64_t bytes = (int64_t)n * (pncu_itemsize(atype) + pncu_itemsize(p->OutType));
    // the reference drops the fast pointer when the output stride is 0 (:882-884, 946-948)
    if (g_Settings.AtopEnabled && p->pUnaryFunc && steps[1] != 0) {
        // integer classifiers (isnan(int) == False everywhere): constant output, see NTE_CLASS_CONST
        const bool is_int = atype <= PNCU_UINT64;
        const int op_class = cat == CAT_TRIG ? NTE_CLASS_HEAVY : ((is_int && const_output) ? NTE_CLASS_CONST : NTE_CLASS_STREAM);
        int rc = nte_unary_class(p->pUnaryFunc, pncu_itemsize(atype), pncu_itemsize(p->OutType), args[0], args[1], (int64_t)n,
                                 (int64_t)steps[0], (int64_t)steps[1], p->MaxThreads,
                                 op_class);
        if (rc == 0) { ledger(cat, true, n, t0, funcop, atype, bytes); return; }
        if (rc != NTE_DECLINED) { raise_cuda_error("unary ufunc loop", rc); return; }
    }
    p->pOldFunc(args, dimensions, steps, innerloop);
    ledger(cat, false, n, t0, funcop, atype, bytes);
}
// reference: AtopUnaryMathFunction, src/pnumpy/_pnumpy.cpp:874-932
static void AtopUnaryMathFunction(char** args, const npy_intp* dimensions, const npy_intp* steps, void* innerloop, int funcop, int atype) {
    const bool classifier = funcop == PNCU_ISNAN || funcop == PNCU_ISINF || funcop == PNCU_ISFINITE || funcop == PNCU_ISNOTNAN ||
                            funcop == PNCU_ISNOTINF || funcop == PNCU_ISNOTFINITE;
    unary_common(&g_UnaryFuncLUT[funcop][atype], CAT_UNARY, args, dimensions, steps, innerloop, atype, funcop, classifier);
}
// reference: AtopTrigMathFunction, src/pnumpy/_pnumpy.cpp:938-996
static void AtopTrigMathFunction(char** args, const npy_intp* dimensions, const npy_intp* steps, void* innerloop, int funcop, int atype) {
    unary_common(&g_TrigFuncLUT[funcop][atype], CAT_TRIG, args, dimensions, steps, innerloop, atype, funcop);
}

// reference: AtopArgMinMaxMathFunction, src/pnumpy/_pnumpy.cpp:1105-1118 (an empty `if (AtopEnabled) {}`
// followed by the old function: the reference never accelerated it; this build does)
static int AtopArgMinMaxMathFunction(void* data, npy_intp n, npy_intp* out_index, void* arr, int kind, int atype) {
    stArgMinMaxFunc* p = &g_ArgMinMaxFuncLUT[kind][atype];
    const uint64_t t0 = ledger_t0();
    const int64_t bytes = (int64_t)n * pncu_itemsize(atype);
    if (g_Settings.AtopEnabled && p->pFunc) {
        int64_t idx = 0;
        int rc = nte_argminmax(kind, atype, data, (int64_t)n, &idx, p->MaxThreads);
        if (rc == 0) { *out_index = (npy_intp)idx; ledger(CAT_ARGMINMAX, true, n, t0, kind, atype, bytes, true); return 0; }
        if (rc != NTE_DECLINED) { raise_cuda_error("arg-reduction", rc); return -1; }
    }
    const int ret = p->pOldFunc(data, n, out_index, arr);
    ledger(CAT_ARGMINMAX, false, n, t0, kind, atype, bytes, true);
    return ret;
}

// =======================================================================================================
// int / int -> float64 true_divide without NumPy's casts (an ADDITION; the 359 reference hooks are
// untouched and these loops are not listed in atop_info()).
//
// np.true_divide has no integer loops (NumPy 2.3: not even 'bb->d'): for int8..int64 operands NumPy casts both
// inputs to float64 in 8 Mi-element buffers on one CPU thread and then runs 'dd->d' -- the only
// thing a replace-loop hook (the reference's, src/atop/ops_binary.cpp:926, or ours) ever sees.  On a
// GPU those casts cost 20x the divide itself.  NumPy 2's public PyUFunc_AddLoopFromSpec lets us add
// exact (intN, intN) -> float64 loops, so the operands reach the dispatcher as they are and the
// kernel (pncu_GetIntTrueDivide: cvt.rn.f64 of both operands, div.rn.f64) produces bit-for-bit what
// NumPy's cast + divide produce.  When atop is disabled or nte declines, the loop does what NumPy
// would have done, with NumPy's own code: PyArray_CastRawArrays into two small float64 buffers and
// the saved 'dd->d' loop -- no arithmetic of ours runs on the CPU.
static int g_ExtraLoops = 0;

template <class T>
static void int_to_double_t(const char* src, npy_intp stride, double* dst, npy_intp n) {
    for (npy_intp i = 0; i < n; ++i) dst[i] = (double)*reinterpret_cast<const T*>(src + i * stride);
}
static void int_to_double(int type_num, const char* src, npy_intp stride, double* dst, npy_intp n) {
    switch (type_num) {
        case NPY_BYTE: int_to_double_t<npy_byte>(src, stride, dst, n); break;
        case NPY_UBYTE: int_to_double_t<npy_ubyte>(src, stride, dst, n); break;
        case NPY_SHORT: int_to_double_t<npy_short>(src, stride, dst, n); break;
        case NPY_USHORT: int_to_double_t<npy_ushort>(src, stride, dst, n); break;
        case NPY_INT: int_to_double_t<npy_int>(src, stride, dst, n); break;
        case NPY_UINT: int_to_double_t<npy_uint>(src, stride, dst, n); break;
        case NPY_LONG: int_to_double_t<npy_long>(src, stride, dst, n); break;
        case NPY_ULONG: int_to_double_t<npy_ulong>(src, stride, dst, n); break;
        case NPY_LONGLONG: int_to_double_t<npy_longlong>(src, stride, dst, n); break;
        default: int_to_double_t<npy_ulonglong>(src, stride, dst, n); break;
    }
}

static int int_true_divide_loop(PyArrayMethod_Context* context, char* const* data, const npy_intp* dimensions,
                                const npy_intp* strides, NpyAuxData*) {
    const int type_num = context->descriptors[0]->type_num;
    const int atype = convert_dtype_to_atop[type_num];
    const npy_intp n = dimensions[0];
    const uint64_t t0 = ledger_t0();
    const int64_t bytes = (int64_t)n * (2 * pncu_itemsize(atype) + 8);
    if (g_Settings.AtopEnabled) {
        pncu_any_two_fn fn = pncu_GetIntTrueDivide(atype);
        int rc = fn ? nte_binary(fn, pncu_itemsize(atype), 8, data[0], data[1], data[2], (int64_t)n, (int64_t)strides[0],
                                 (int64_t)strides[1], (int64_t)strides[2], 3)
                    : NTE_DECLINED;
        if (rc == 0) { ledger(CAT_BINARY, true, n, t0, PNCU_DIV, atype, bytes); return 0; }
        if (rc != NTE_DECLINED) { raise_cuda_error("int true_divide loop", rc); return -1; }
    }
    // What NumPy itself does for this call: cast both operands to float64, then its own 'dd->d' loop (which also
    // raises the IEEE flags NumPy's divide-by-zero / invalid warnings come from).  The casts are plain C
    // conversions (exact for every integer up to 2^53, round-to-nearest beyond: the hardware's cvtsi2sd, which is
    // what NumPy's cast loops compile to), done chunk by chunk into two small stack buffers: no Python API, so
    // the GIL stays wherever NumPy left it, and nothing is allocated.
    PyUFuncGenericFunction dd = g_UFuncLUT[PNCU_DIV][PNCU_DOUBLE].pOldFunc;
    const npy_intp CHUNK = 2048;
    double buf[2][CHUNK];
    for (npy_intp off = 0; off < n; off += CHUNK) {
        const npy_intp cnt = (n - off) < CHUNK ? (n - off) : CHUNK;
        for (int k = 0; k < 2; ++k) int_to_double(type_num, data[k] + off * strides[k], strides[k], buf[k], cnt);
        char* a3[3] = {(char*)buf[0], (char*)buf[1], data[2] + off * strides[2]};
        const npy_intp dims[1] = {cnt};
        const npy_intp steps[3] = {8, 8, strides[2]};
        dd(a3, dims, steps, NULL);
    }
    ledger(CAT_BINARY, false, n, t0, PNCU_DIV, atype, bytes);
    return 0;
}

static int register_int_true_divide(PyObject* numpy_module) {
    if (!g_UFuncLUT[PNCU_DIV][PNCU_DOUBLE].pOldFunc) return 0;
    PyObject* ufunc = PyObject_GetAttrString(numpy_module, "true_divide");
    if (!ufunc) return -1;
    // (NumPy 2 has no 'bb->d' loop any more: int8 is cast to float64 like the others)
    PyArray_DTypeMeta* ints[] = {&PyArray_ByteDType, &PyArray_UByteDType, &PyArray_ShortDType, &PyArray_UShortDType, &PyArray_IntDType,
                                 &PyArray_UIntDType, &PyArray_LongDType, &PyArray_ULongDType, &PyArray_LongLongDType,
                                 &PyArray_ULongLongDType};
    for (PyArray_DTypeMeta* dt : ints) {
        PyArray_DTypeMeta* dtypes[3] = {dt, dt, &PyArray_DoubleDType};
        PyType_Slot slots[] = {{NPY_METH_strided_loop, (void*)int_true_divide_loop}, {0, NULL}};
        PyArrayMethod_Spec spec;
        memset(&spec, 0, sizeof(spec));
        spec.name = "pnumpy_b200_int_true_divide";
        spec.nin = 2;
        spec.nout = 1;
        spec.casting = NPY_NO_CASTING;
        spec.flags = (NPY_ARRAYMETHOD_FLAGS)0;
        spec.dtypes = dtypes;
        spec.slots = slots;
        if (PyUFunc_AddLoopFromSpec(ufunc, &spec) < 0) {
            PyErr_Clear();  // e.g. a loop for this pair already exists: leave NumPy's behaviour alone
            continue;
        }
        g_ExtraLoops++;
    }
    Py_DECREF(ufunc);
    return 0;
}

// =======================================================================================================
// registration.  reference: newinit, src/pnumpy/_pnumpy.cpp:1287-1676
static void AddToDict(const char* ufunc_name, int dtype, void* pstUFunc) {
    PyObject* key = Py_BuildValue("(si)", ufunc_name, dtype);
    PyObject* val = PyLong_FromLongLong((long long)(intptr_t)pstUFunc);
    PyDict_SetItem(gpUnaryDict, key, val);
    Py_DECREF(key);
    Py_DECREF(val);
}
static stUFunc* GetFromDict(const char* ufunc_name, int dtype) {
    if (!gpUnaryDict) return NULL;
    PyObject* key = Py_BuildValue("(si)", ufunc_name, dtype);
    PyObject* val = PyDict_GetItem(gpUnaryDict, key);  // borrowed
    Py_DECREF(key);
    return val ? (stUFunc*)(intptr_t)PyLong_AsLongLong(val) : NULL;
}

// "Check for problem with two int32 depending on OS" (:1357-1362): keep the input's own type number
static int fix_out_dtype(int in_dtype, int out_atype) {
    int out_dtype = convert_atop_to_dtype[out_atype];
    if (in_dtype >= 5 && in_dtype <= 10 && out_dtype >= 5 && out_dtype <= 10) out_dtype = in_dtype;
    return out_dtype;
}

static PyObject* newinit(PyObject* self, PyObject* args, PyObject* kwargs) {
    // NPY_BOOL .. NPY_USHORT, NPY_INT, NPY_UINT, NPY_LONG, NPY_ULONG, NPY_LONGLONG, NPY_ULONGLONG, NPY_FLOAT, NPY_DOUBLE
    static const int dtypes[] = {NPY_BOOL, NPY_INT8, NPY_UINT8, NPY_INT16, NPY_UINT16, 5, 6, 7, 8, 9, 10, 11, 12};
    if (g_Settings.Initialized) Py_RETURN_NONE;  // idempotent (:1294)

    // the analogue of the reference's AVX2 gate (src/pnumpy/__init__.py:48-53): no sm_100 GPU, no pnumpy
    int rc = nte_init();
    if (rc != 0)
        return PyErr_Format(PyExc_ImportError, "pnumpy_b200 requires an NVIDIA B200 (sm_100) GPU: %s", pncu_last_error());
    {
        char one[256];
        int n = nte_device_count(), off = 0;
        for (int d = 0; d < n && off < (int)sizeof(g_DeviceString) - 1; ++d) {
            if (pncu_device_info(d, one, sizeof(one)) != 0) continue;
            off += snprintf(g_DeviceString + off, sizeof(g_DeviceString) - off, "%s[%d] %s", d ? "; " : "**GPU: ", d, one);
        }
    }

    memset(g_UFuncLUT, 0, sizeof(g_UFuncLUT));
    memset(g_CompFuncLUT, 0, sizeof(g_CompFuncLUT));
    memset(g_UnaryFuncLUT, 0, sizeof(g_UnaryFuncLUT));
    memset(g_TrigFuncLUT, 0, sizeof(g_TrigFuncLUT));
    memset(g_ArgMinMaxFuncLUT, 0, sizeof(g_ArgMinMaxFuncLUT));

    PyObject* numpy_module = PyImport_ImportModule("numpy");
    if (!numpy_module) return NULL;

    // np.setbufsize(8192*1024): casting / buffered calls then deliver <= 8Mi-element loops (:1306-1314)
    PyObject* r = PyObject_CallMethod(numpy_module, "setbufsize", "L", (long long)8192 * 1024);
    if (!r) PyErr_Clear();
    Py_XDECREF(r);

    gpUnaryDict = PyDict_New();
    const int num_dtypes = COUNT_OF(dtypes);

    // ---- binary (:1320-1380)
    for (int i = 0; i < COUNT_OF(gBinaryMapping); i++) {
        const char* name = gBinaryMapping[i].str_ufunc_name;
        const int atop = gBinaryMapping[i].atop_op;
        PyObject* ufunc = PyObject_GetAttrString(numpy_module, name);
        if (!ufunc) return PyErr_Format(PyExc_TypeError, "func %s must be the name of a ufunc", name);
        for (int jx = 0; jx < num_dtypes; jx++) {
            const int dtype = dtypes[jx];
            const int atype = convert_dtype_to_atop[dtype];
            int out_atype = -1;
            pncu_any_two_fn fn = pncu_GetSimpleMathOpFast(atop, atype, atype, &out_atype);
            pncu_reduce_fn red = pncu_GetReduceMathOpFast(atop, atype);
            if (out_atype == -1) continue;  // the pair is not hooked at all (:1349)
            int signature[3] = {dtype, dtype, fix_out_dtype(dtype, out_atype)};
            PyUFuncGenericFunction oldFunc;
            int ret = PyUFunc_ReplaceLoopBySignature((PyUFuncObject*)ufunc, g_UFuncGenericLUT[atop][atype], signature, &oldFunc);
            if (ret < 0)
                return PyErr_Format(PyExc_TypeError, "Binary failed with %d. func %s must be the name of a ufunc.  atop:%d   atype:%d  sigs:%d, %d, %d",
                                    ret, name, atop, atype, signature[0], signature[1], signature[2]);
            stUFunc* p = &g_UFuncLUT[atop][atype];
            p->pOldFunc = oldFunc;
            p->pBinaryFunc = fn;
            p->pReduceFunc = red;
            p->MaxThreads = 3;
            p->OutType = out_atype;
            AddToDict(name, dtype, p);
        }
        Py_DECREF(ufunc);
    }
    // ---- compare (:1383-1432): always hooked, out is bool
    for (int i = 0; i < COUNT_OF(gCompareMapping); i++) {
        const char* name = gCompareMapping[i].str_ufunc_name;
        const int atop = gCompareMapping[i].atop_op;
        PyObject* ufunc = PyObject_GetAttrString(numpy_module, name);
        if (!ufunc) return PyErr_Format(PyExc_TypeError, "func %s must be the name of a ufunc", name);
        for (int jx = 0; jx < num_dtypes; jx++) {
            const int dtype = dtypes[jx];
            const int atype = convert_dtype_to_atop[dtype];
            int out_atype = PNCU_BOOL;
            pncu_any_two_fn fn = pncu_GetComparisonOpFast(atop, atype, atype, &out_atype);
            int signature[3] = {dtype, dtype, fix_out_dtype(dtype, out_atype)};
            PyUFuncGenericFunction oldFunc;
            int ret = PyUFunc_ReplaceLoopBySignature((PyUFuncObject*)ufunc, g_UFuncCompareLUT[atop][atype], signature, &oldFunc);
            if (ret < 0)
                return PyErr_Format(PyExc_TypeError, "Comparison failed with %d. func %s must be the name of a ufunc.  atop:%d   atype:%d  sigs:%d, %d, %d",
                                    ret, name, atop, atype, signature[0], signature[1], signature[2]);
            stUFunc* p = &g_CompFuncLUT[atop][atype];
            p->pOldFunc = oldFunc;
            p->pBinaryFunc = fn;
            p->MaxThreads = 3;
            p->OutType = PNCU_BOOL;
            AddToDict(name, dtype, p);
        }
        Py_DECREF(ufunc);
    }
    // ---- unary (:1435-1488)
    for (int i = 0; i < COUNT_OF(gUnaryMapping); i++) {
        const char* name = gUnaryMapping[i].str_ufunc_name;
        const int atop = gUnaryMapping[i].atop_op;
        PyObject* ufunc = PyObject_GetAttrString(numpy_module, name);
        if (!ufunc) return PyErr_Format(PyExc_TypeError, "func %s must be the name of a ufunc", name);
        for (int jx = 0; jx < num_dtypes; jx++) {
            const int dtype = dtypes[jx];
            const int atype = convert_dtype_to_atop[dtype];
            int out_atype = -1;
            pncu_unary_fn fn = pncu_GetUnaryOpFast(atop, atype, &out_atype);
            if (out_atype == -1) continue;
            int signature[3] = {dtype, fix_out_dtype(dtype, out_atype), dtype};
            PyUFuncGenericFunction oldFunc;
            int ret = PyUFunc_ReplaceLoopBySignature((PyUFuncObject*)ufunc, g_UFuncUnaryLUT[atop][atype], signature, &oldFunc);
            if (ret < 0)
                return PyErr_Format(PyExc_TypeError, "Unary failed with %d. func %s must be the name of a ufunc.  atop:%d   atype:%d  sigs:%d, %d, %d",
                                    ret, name, atop, atype, signature[0], signature[1], signature[2]);
            stUFunc* p = &g_UnaryFuncLUT[atop][atype];
            p->pOldFunc = oldFunc;
            p->pUnaryFunc = fn;
            p->MaxThreads = 3;
            p->OutType = out_atype;
            AddToDict(name, dtype, p);
        }
        Py_DECREF(ufunc);
    }
    // ---- trig / log / exp (:1492-1547): float32 and float64, hooked even without a body, MaxThreads 7
    static const int trig_dtypes[] = {NPY_FLOAT32, NPY_FLOAT64};
    for (int i = 0; i < COUNT_OF(gTrigMapping); i++) {
        const char* name = gTrigMapping[i].str_ufunc_name;
        const int atop = gTrigMapping[i].atop_op;
        PyObject* ufunc = PyObject_GetAttrString(numpy_module, name);
        if (!ufunc) return PyErr_Format(PyExc_TypeError, "func %s must be the name of a ufunc", name);
        for (int jx = 0; jx < COUNT_OF(trig_dtypes); jx++) {
            const int dtype = trig_dtypes[jx];
            const int atype = convert_dtype_to_atop[dtype];
            int out_atype = atype;
            pncu_unary_fn fn = pncu_GetTrigOpFast(atop, atype, &out_atype);
            int signature[3] = {dtype, convert_atop_to_dtype[out_atype], dtype};
            PyUFuncGenericFunction oldFunc;
            int ret = PyUFunc_ReplaceLoopBySignature((PyUFuncObject*)ufunc, g_UFuncTrigLUT[atop][atype], signature, &oldFunc);
            if (ret < 0)
                return PyErr_Format(PyExc_TypeError, "Trig failed with %d. func %s must be the name of a ufunc.  atop:%d   atype:%d  sigs:%d, %d, %d",
                                    ret, name, atop, atype, signature[0], signature[1], signature[2]);
            stUFunc* p = &g_TrigFuncLUT[atop][atype];
            p->pOldFunc = oldFunc;
            p->pUnaryFunc = fn;  // NULL allowed: the old loop runs
            p->MaxThreads = 7;
            p->OutType = out_atype;
            AddToDict(name, dtype, p);
        }
        Py_DECREF(ufunc);
    }
    // ---- PyArray_ArrFuncs: argmax / argmin (:1600-1614).  sort, argsort and fill are out of scope
    //  and are left to NumPy.
    for (int jx = 0; jx < num_dtypes; jx++) {
        const int dtype = dtypes[jx];
        const int atype = convert_dtype_to_atop[dtype];
        PyArray_Descr* descr = PyArray_DescrFromType(dtype);
        if (!descr) { PyErr_Clear(); continue; }
        PyArray_ArrFuncs* f = PyDataType_GetArrFuncs(descr);
        if (f) {
            stArgMinMaxFunc* p = &g_ArgMinMaxFuncLUT[0][atype];
            if (f->argmax && f->argmax != (PyArray_ArgFunc*)g_UFuncArgMinMaxLUT[0][atype]) {
                p->pOldFunc = (ArgMinMaxFunc)f->argmax;
                p->pFunc = pncu_GetArgMinMaxOpFast(PNCU_ARGMAX, atype);
                p->MaxThreads = 3;
                f->argmax = (PyArray_ArgFunc*)g_UFuncArgMinMaxLUT[0][atype];
            }
            p = &g_ArgMinMaxFuncLUT[1][atype];
            if (f->argmin && f->argmin != (PyArray_ArgFunc*)g_UFuncArgMinMaxLUT[1][atype]) {
                p->pOldFunc = (ArgMinMaxFunc)f->argmin;
                p->pFunc = pncu_GetArgMinMaxOpFast(PNCU_ARGMIN, atype);
                p->MaxThreads = 3;
                f->argmin = (PyArray_ArgFunc*)g_UFuncArgMinMaxLUT[1][atype];
            }
        }
        Py_DECREF(descr);
    }
    if (register_int_true_divide(numpy_module) < 0) return NULL;
    Py_DECREF(numpy_module);
    g_Settings.Initialized = 1;
    Py_RETURN_NONE;
}

// =======================================================================================================
// control surface.  reference: src/pnumpy/_pnumpy.cpp:1766-1966
static PyObject* atop_enable(PyObject*, PyObject*) { g_Settings.AtopEnabled = 1; Py_RETURN_NONE; }
static PyObject* atop_disable(PyObject*, PyObject*) { g_Settings.AtopEnabled = 0; Py_RETURN_NONE; }
static PyObject* atop_isenabled(PyObject*, PyObject*) { return PyBool_FromLong(g_Settings.AtopEnabled); }

// atop_info() -> the live dict; atop_info(name, dtype_num) -> per-entry dict or False (:1791-1818)
static PyObject* atop_info(PyObject*, PyObject* args) {
    if (!g_Settings.Initialized) Py_RETURN_FALSE;
    if (PyTuple_Size(args) == 0) {
        Py_INCREF(gpUnaryDict);
        return gpUnaryDict;
    }
    const char* uname = NULL;
    int dtype = 0;
    if (!PyArg_ParseTuple(args, "si:atop_info", &uname, &dtype)) return NULL;
    stUFunc* p = GetFromDict(uname, dtype);
    if (!p) Py_RETURN_FALSE;
    // a compare / binary entry keeps its launcher in the pBinaryFunc/pUnaryFunc union, as in the reference
    return Py_BuildValue("{s:i,s:i,s:L,s:L,s:L,s:L}", "MaxThreads", (int)p->MaxThreads, "MinElementsToThread",
                         (int)p->MinElementsToThread, "pBinaryFunc", (long long)(intptr_t)p->pBinaryFunc, "pOldFunc",
                         (long long)(intptr_t)p->pOldFunc, "pReduceFunc", (long long)(intptr_t)p->pReduceFunc,
                         "pUnaryFunc", (long long)(intptr_t)p->pUnaryFunc);
}

// atop_setworkers(name, dtype_num, n): 0 <= n < 64, returns the previous value or False, only while
// atop is enabled (:1825-1846)
static PyObject* atop_setworkers(PyObject*, PyObject* args) {
    if (g_Settings.AtopEnabled) {
        const char* uname = NULL;
        int dtype = 0, workers = 0;
        if (!PyArg_ParseTuple(args, "sii:atop_setworkers", &uname, &dtype, &workers)) return NULL;
        stUFunc* p = GetFromDict(uname, dtype);
        if (p && workers >= 0 && workers < 64) {
            int32_t prev = p->MaxThreads;
            p->MaxThreads = workers;
            return PyLong_FromLong(prev);
        }
    }
    Py_RETURN_FALSE;
}

static PyObject* thread_enable(PyObject*, PyObject*) { nte_set_threading(1); Py_RETURN_NONE; }
static PyObject* thread_disable(PyObject*, PyObject*) { nte_set_threading(0); Py_RETURN_NONE; }
static PyObject* thread_isenabled(PyObject*, PyObject*) { return PyBool_FromLong(nte_get_threading()); }
static PyObject* thread_setworkers(PyObject*, PyObject* args) {
    int workers = 0;
    if (!PyArg_ParseTuple(args, "i:thread_setworkers", &workers)) return NULL;
    return PyLong_FromLong(nte_set_workers(workers));  // clamps, returns the previous value
}
static PyObject* thread_getworkers(PyObject*, PyObject*) { return PyLong_FromLong(nte_get_workers()); }
static PyObject* thread_zigzag(PyObject*, PyObject*) {
    // toggles 0 <-> 3 and reports the new state (:1953-1966); the value has no effect on a GPU
    g_Settings.ZigZag = g_Settings.ZigZag ? 0 : 3;
    return PyBool_FromLong(g_Settings.ZigZag != 0);
}

static PyObject* cpustring(PyObject*, PyObject*) {
    // the reference returns CMathWorker::CPUString ("**CPU: <brand> AVX2:1 ..."); here the devices
    if (!g_DeviceString[0]) Py_RETURN_NONE;
    return PyUnicode_FromString(g_DeviceString);
}

static PyObject* ledger_enable(PyObject*, PyObject*) { g_Settings.LedgerEnabled = 1; Py_RETURN_NONE; }
static PyObject* ledger_disable(PyObject*, PyObject*) { g_Settings.LedgerEnabled = 0; Py_RETURN_NONE; }
static PyObject* ledger_isenabled(PyObject*, PyObject*) { return PyBool_FromLong(g_Settings.LedgerEnabled); }
static const char* kLedgerCatNames[5] = {"Binary", "Unary", "Compare", "TrigLog", "ArgMinMax"};
static PyObject* ledger_info(PyObject*, PyObject*) {
    PyObject* d = PyDict_New();
    std::lock_guard<std::mutex> lk(g_LedgerMu);
    for (int c = 0; c < 5; ++c) {
        PyObject* v = Py_BuildValue("{s:L,s:L,s:L}", "gpu_calls", (long long)g_Ledger[c].gpu_calls, "numpy_calls",
                                    (long long)g_Ledger[c].old_calls, "elements", (long long)g_Ledger[c].elements);
        PyDict_SetItemString(d, kLedgerCatNames[c], v);
        Py_DECREF(v);
    }
    return d;
}
// ledger_records(last=None) -> the most recent per-call records, oldest first (the reference's ledger ring,
// src/pnumpy/ledger.cpp:144-272: category, op, type, length, start/end ticks), plus where the call ran
static PyObject* ledger_records(PyObject*, PyObject* args) {
    long long last = kLedgerRing;
    if (!PyArg_ParseTuple(args, "|L:ledger_records", &last)) return NULL;
    std::vector<LedgerRecord> recs;
    {
        std::lock_guard<std::mutex> lk(g_LedgerMu);
        uint64_t have = g_LedgerHead < (uint64_t)kLedgerRing ? g_LedgerHead : (uint64_t)kLedgerRing;
        if (last >= 0 && (uint64_t)last < have) have = (uint64_t)last;
        for (uint64_t i = g_LedgerHead - have; i < g_LedgerHead; ++i) recs.push_back(g_LedgerRing[i % kLedgerRing]);
    }
    PyObject* list = PyList_New((Py_ssize_t)recs.size());
    if (!list) return NULL;
    for (size_t i = 0; i < recs.size(); ++i) {
        const LedgerRecord& r = recs[i];
        PyObject* v = Py_BuildValue("{s:s,s:i,s:i,s:L,s:L,s:K,s:K,s:K,s:O,s:O}", "category", kLedgerCatNames[r.cat], "op", (int)r.funcop,
                                    "atype", (int)r.atype, "elements", (long long)r.n, "bytes", (long long)r.bytes, "start_tsc",
                                    (unsigned long long)r.t0, "end_tsc", (unsigned long long)r.t1, "ticks",
                                    (unsigned long long)(r.t1 - r.t0), "gpu", r.gpu ? Py_True : Py_False, "reduce",
                                    r.reduce ? Py_True : Py_False);
        if (!v) { Py_DECREF(list); return NULL; }
        PyList_SET_ITEM(list, (Py_ssize_t)i, v);
    }
    return list;
}
static PyObject* ledger_clear(PyObject*, PyObject*) {
    std::lock_guard<std::mutex> lk(g_LedgerMu);
    memset(g_Ledger, 0, sizeof(g_Ledger));
    g_LedgerHead = 0;
    Py_RETURN_NONE;
}

// ---- recycler: NEP-49 data allocator that gives large ndarrays memory the GPU can reach, and recycles freed
// blocks (the reference's roadmap for its placeholder recycler, src/pnumpy/__init__.py:18-21,
// src/pnumpy/recycler.cpp, doc_src/source/roadmap.rst:27-31).  Two kinds:
//   "pinned"   page-locked host memory: nte DMAs such operands in place (no staging copy); every call still
//              crosses PCIe.
//   "managed"  CUDA managed memory: the arrays NumPy creates -- np.empty, ufunc outputs, temporaries of
//              a*b+c -- are DEVICE-CAPABLE from birth.  A hooked call whose operands are managed runs in place
//              on the GPU at any size; intermediates never cross PCIe; pages migrate to the host only when host
//              code touches them.  This is what makes residency reachable through the unmodified NumPy API.
// Blocks are recycled by exact size (NumPy asks for the same sizes over and over); a recycled managed block
// keeps its pages where the last kernel left them. ------------------------------------------------------------
enum { RK_PINNED = 0, RK_MANAGED = 1 };
static size_t g_RecyclerThreshold[2] = {(size_t)1 << 20, (size_t)256 << 10};
static int g_RecyclerKind = RK_PINNED;
static std::mutex g_PoolMu;
struct PoolBlock { void* p; uint64_t seq; };
static std::multimap<size_t, PoolBlock> g_Pool[2];
// idle bytes kept per kind: pinned blocks are host RAM (8 GiB); managed blocks are HBM the next temporary of the same
// size will reuse (64 GiB of 180 per GPU; a 2^30-element float32 chain alone cycles through 8 GiB).  When a freed block
// does not fit, the OLDEST idle blocks make room for it (the sizes a program asks for change over time).
static size_t g_PoolBytes[2] = {0, 0}, g_PoolCap[2] = {(size_t)8 << 30, (size_t)64 << 30};
static uint64_t g_PoolSeq = 0;
static int64_t g_PoolHits = 0, g_PoolMisses = 0, g_BlocksLive = 0;

struct LiveBlock { size_t size; int kind; };
static std::map<void*, LiveBlock> g_Live;  // blocks currently owned by ndarrays

static void* pool_get(size_t size, int kind) {
    void* p = NULL;
    {
        std::lock_guard<std::mutex> lk(g_PoolMu);
        auto it = g_Pool[kind].find(size);
        if (it != g_Pool[kind].end()) {
            p = it->second.p;
            g_Pool[kind].erase(it);
            g_PoolBytes[kind] -= size;
            g_PoolHits++;
        } else {
            g_PoolMisses++;
        }
    }
    if (!p) p = kind == RK_MANAGED ? nte_alloc_managed(size) : nte_alloc_pinned(size);
    if (p) {
        std::lock_guard<std::mutex> lk(g_PoolMu);
        g_Live[p] = LiveBlock{size, kind};
        g_BlocksLive++;
    }
    return p;
}
// true (and the block is retired) if `p` is one of ours
static bool pool_put(void* p) {
    LiveBlock b;
    std::vector<void*> evicted;
    bool pooled = false;
    {
        std::lock_guard<std::mutex> lk(g_PoolMu);
        auto it = g_Live.find(p);
        if (it == g_Live.end()) return false;
        b = it->second;
        g_Live.erase(it);
        g_BlocksLive--;
        if (!nte_is_forked_child() && b.size <= g_PoolCap[b.kind]) {
            while (g_PoolBytes[b.kind] + b.size > g_PoolCap[b.kind] && !g_Pool[b.kind].empty()) {
                auto oldest = g_Pool[b.kind].begin();
                for (auto it = g_Pool[b.kind].begin(); it != g_Pool[b.kind].end(); ++it)
                    if (it->second.seq < oldest->second.seq) oldest = it;
                evicted.push_back(oldest->second.p);
                g_PoolBytes[b.kind] -= oldest->first;
                g_Pool[b.kind].erase(oldest);
            }
            g_Pool[b.kind].insert(std::make_pair(b.size, PoolBlock{p, ++g_PoolSeq}));
            g_PoolBytes[b.kind] += b.size;
            pooled = true;
        }
    }
    for (void* q : evicted) { if (b.kind == RK_MANAGED) nte_free_managed(q); else nte_free_pinned(q); }
    if (pooled) return true;
    if (b.kind == RK_MANAGED) nte_free_managed(p); else nte_free_pinned(p);
    return true;
}
static bool live_block_of(void* p, LiveBlock* b) {
    std::lock_guard<std::mutex> lk(g_PoolMu);
    auto it = g_Live.find(p);
    if (it == g_Live.end()) return false;
    *b = it->second;
    return true;
}
static void* rc_malloc(void*, size_t size) {
    const int kind = g_RecyclerKind;
    if (size >= g_RecyclerThreshold[kind]) { void* p = pool_get(size, kind); if (p) return p; }
    return malloc(size ? size : 1);
}
static void* rc_calloc(void*, size_t nelem, size_t elsize) {
    const size_t size = nelem * elsize;
    const int kind = g_RecyclerKind;
    if (size >= g_RecyclerThreshold[kind]) {
        void* p = pool_get(size, kind);
        if (p) {
            // managed: zero on the GPU, so np.zeros() does not drag the pages to the host
            if (kind != RK_MANAGED || nte_zero_managed(p, size) != 0) memset(p, 0, size);
            return p;
        }
    }
    return calloc(nelem ? nelem : 1, elsize ? elsize : 1);
}
static void rc_free(void*, void* ptr, size_t) {
    if (!ptr) return;
    if (!pool_put(ptr)) free(ptr);
}
static void* rc_realloc(void* ctx, void* ptr, size_t new_size) {
    if (!ptr) return rc_malloc(ctx, new_size);
    LiveBlock b;
    if (!live_block_of(ptr, &b)) {
        // a malloc'd block: plain realloc keeps the data valid; grown past the threshold it simply stays
        // pageable (staged through nte's slabs like any other host array)
        return realloc(ptr, new_size ? new_size : 1);
    }
    void* q = rc_malloc(ctx, new_size);
    if (!q) return NULL;
    memcpy(q, ptr, b.size < new_size ? b.size : new_size);
    pool_put(ptr);
    return q;
}
static PyDataMem_Handler g_RecyclerHandler = {"pnumpy_b200_recycler", 1, {NULL, rc_malloc, rc_calloc, rc_realloc, rc_free}};
static PyObject* g_OldHandler = NULL;

// recycler_enable(kind="pinned", threshold=None): the reference's recycler_enable() takes no arguments and so
// keeps working; kind="managed" makes large ndarrays device-capable (see above).  Env PNUMPY_B200_RECYCLER sets
// the default kind.
static PyObject* recycler_enable(PyObject*, PyObject* args, PyObject* kwargs) {
    static const char* kwlist[] = {"kind", "threshold", NULL};
    const char* kind = NULL;
    PyObject* threshold = Py_None;
    if (!PyArg_ParseTupleAndKeywords(args, kwargs, "|zO:recycler_enable", (char**)kwlist, &kind, &threshold)) return NULL;
    if (!g_Settings.Initialized) { PyErr_SetString(PyExc_RuntimeError, "call initialize() first"); return NULL; }
    if (!kind) kind = getenv("PNUMPY_B200_RECYCLER");
    int k = RK_PINNED;
    if (kind && *kind) {
        if (!strcmp(kind, "managed")) k = RK_MANAGED;
        else if (strcmp(kind, "pinned")) { PyErr_SetString(PyExc_ValueError, "kind must be 'pinned' or 'managed'"); return NULL; }
    }
    if (threshold != Py_None) {
        const long long t = PyLong_AsLongLong(threshold);
        if (t < 0) { if (!PyErr_Occurred()) PyErr_SetString(PyExc_ValueError, "threshold must be >= 0"); return NULL; }
        g_RecyclerThreshold[k] = (size_t)(t < 4096 ? 4096 : t);
    }
    g_RecyclerKind = k;
    const char* cap_gib = getenv("PNUMPY_B200_POOL_GIB");   // idle bytes the pool of this kind may keep
    if (cap_gib && *cap_gib && atoll(cap_gib) >= 0) g_PoolCap[k] = (size_t)atoll(cap_gib) << 30;
    if (!g_Settings.RecyclerEnabled) {
        PyObject* cap = PyCapsule_New(&g_RecyclerHandler, "mem_handler", NULL);
        if (!cap) return NULL;
        PyObject* old = PyDataMem_SetHandler(cap);
        Py_DECREF(cap);
        if (!old) return NULL;
        Py_XDECREF(g_OldHandler);
        g_OldHandler = old;
        g_Settings.RecyclerEnabled = 1;
    }
    Py_RETURN_NONE;
}
static PyObject* recycler_disable(PyObject*, PyObject*) {
    if (g_Settings.RecyclerEnabled) {
        PyObject* old = PyDataMem_SetHandler(g_OldHandler);  // arrays keep the handler they were born with
        Py_XDECREF(old);
        Py_CLEAR(g_OldHandler);
        g_Settings.RecyclerEnabled = 0;
    }
    Py_RETURN_NONE;
}
static PyObject* recycler_isenabled(PyObject*, PyObject*) { return PyBool_FromLong(g_Settings.RecyclerEnabled); }
static PyObject* recycler_info(PyObject*, PyObject*) {
    std::lock_guard<std::mutex> lk(g_PoolMu);
    return Py_BuildValue("{s:s,s:L,s:L,s:L,s:L,s:L,s:L,s:L,s:L,s:L}", "kind", g_RecyclerKind == RK_MANAGED ? "managed" : "pinned",
                         "pooled_blocks", (long long)(g_Pool[0].size() + g_Pool[1].size()), "pooled_bytes",
                         (long long)(g_PoolBytes[0] + g_PoolBytes[1]), "pooled_managed_bytes", (long long)g_PoolBytes[1],
                         "hits", (long long)g_PoolHits, "misses", (long long)g_PoolMisses,
                         "live_blocks", (long long)g_BlocksLive, "live_pinned_blocks", (long long)g_BlocksLive,
                         "threshold_bytes", (long long)g_RecyclerThreshold[g_RecyclerKind],
                         "pool_cap_bytes", (long long)g_PoolCap[g_RecyclerKind]);
}
// give pooled (free) blocks back to CUDA; live arrays are untouched
static PyObject* recycler_trim(PyObject*, PyObject*) {
    std::vector<std::pair<void*, int>> victims;
    {
        std::lock_guard<std::mutex> lk(g_PoolMu);
        for (int k = 0; k < 2; ++k) {
            for (auto& kv : g_Pool[k]) victims.push_back(std::make_pair(kv.second.p, k));
            g_Pool[k].clear();
            g_PoolBytes[k] = 0;
        }
    }
    for (auto& v : victims) { if (v.second == RK_MANAGED) nte_free_managed(v.first); else nte_free_pinned(v.first); }
    return PyLong_FromSize_t(victims.size());
}

// ---- timers.  reference: src/pnumpy/ledger.cpp:91-114 ---------------------------------------------------
static inline uint64_t rdtsc_now(void) {
#if defined(__x86_64__) || defined(__i386__)
    unsigned hi, lo;
    __asm__ __volatile__("rdtsc" : "=a"(lo), "=d"(hi));
    return ((uint64_t)lo) | (((uint64_t)hi) << 32);
#else
    struct timespec x;
    clock_gettime(CLOCK_MONOTONIC, &x);
    return (uint64_t)x.tv_sec * 1000000000ull + (uint64_t)x.tv_nsec;
#endif
}
static PyObject* timer_gettsc(PyObject*, PyObject*) { return PyLong_FromUnsignedLongLong(rdtsc_now()); }
static PyObject* timer_getutc(PyObject*, PyObject*) {
    struct timespec x;
    clock_gettime(CLOCK_REALTIME, &x);
    return PyLong_FromLongLong((long long)x.tv_sec * 1000000000LL + x.tv_nsec);
}

// ---- additions (new keys only; nothing above changes meaning) ----------------------------------------------
static PyObject* cuda_info(PyObject*, PyObject*) {
    nte_stats s;
    nte_get_stats(&s);
    return Py_BuildValue("{s:i,s:s,s:L,s:L,s:L,s:L,s:L,s:L,s:L,s:L,s:L,s:L,s:i,s:i,s:i,s:L,s:i}", "devices", nte_device_count(), "device_string",
                         g_DeviceString, "min_elems", (long long)nte_get_min_elems(), "calls", (long long)s.calls, "declined",
                         (long long)s.declined, "launches", (long long)s.launches, "library_launches", (long long)pncu_launch_count(),
                         "h2d_bytes", (long long)s.h2d_bytes, "d2h_bytes", (long long)s.d2h_bytes, "staged_bytes",
                         (long long)s.staged_bytes, "lanes_last", (long long)s.lanes_last, "gpus_last", (long long)s.gpus_last,
                         "workers", nte_get_workers(), "threading", nte_get_threading(), "extra_loops", g_ExtraLoops,
                         "prefetches", (long long)s.prefetches, "forked_child", nte_is_forked_child());
}
// cuda_trace(on=None): switch the dispatcher's per-call trace on / off; returns the last call's eight stations (ns)
static PyObject* cuda_trace(PyObject*, PyObject* args) {
    PyObject* on = Py_None;
    if (!PyArg_ParseTuple(args, "|O:cuda_trace", &on)) return NULL;
    if (on != Py_None) nte_set_trace(PyObject_IsTrue(on));
    int64_t t[8];
    nte_get_trace(t);
    return Py_BuildValue("(LLLLLLLL)", (long long)t[0], (long long)t[1], (long long)t[2], (long long)t[3], (long long)t[4],
                         (long long)t[5], (long long)t[6], (long long)t[7]);
}
// cuda_inject_fault(n): the n-th GPU dispatch from now fails like a bad launch (error-path tests); 0 disarms
static PyObject* cuda_inject_fault(PyObject*, PyObject* args) {
    long long n = 0;
    if (!PyArg_ParseTuple(args, "L:cuda_inject_fault", &n)) return NULL;
    return PyLong_FromLongLong(nte_inject_fault(n));
}
static PyObject* cuda_reset_stats(PyObject*, PyObject*) { nte_reset_stats(); Py_RETURN_NONE; }
static PyObject* cuda_setthreshold(PyObject*, PyObject* args) {
    long long n = 0;
    if (!PyArg_ParseTuple(args, "L:cuda_setthreshold", &n)) return NULL;
    return PyLong_FromLongLong(nte_set_min_elems(n));
}
static PyObject* cuda_getthreshold(PyObject*, PyObject*) { return PyLong_FromLongLong(nte_get_min_elems()); }
static PyObject* cuda_setslab(PyObject*, PyObject* args) {
    long long n = 0;
    if (!PyArg_ParseTuple(args, "L:cuda_setslab", &n)) return NULL;
    return PyLong_FromLongLong(nte_set_slab_bytes(n));
}

// pinned_empty(nbytes) -> a writable buffer object over pinned host memory (freed with the object)
static void pinned_capsule_free(PyObject* cap) { nte_free_pinned(PyCapsule_GetPointer(cap, "pnumpy_b200.pinned")); }
static PyObject* pinned_alloc(PyObject*, PyObject* args) {
    long long nbytes = 0;
    if (!PyArg_ParseTuple(args, "L:pinned_alloc", &nbytes)) return NULL;
    if (nbytes < 0) { PyErr_SetString(PyExc_ValueError, "negative size"); return NULL; }
    void* p = nte_alloc_pinned((size_t)nbytes);
    if (!p) return PyErr_Format(PyExc_MemoryError, "cudaHostAlloc(%lld) failed: %s", nbytes, pncu_last_error());
    PyObject* cap = PyCapsule_New(p, "pnumpy_b200.pinned", pinned_capsule_free);
    if (!cap) { nte_free_pinned(p); return NULL; }
    return Py_BuildValue("(NL)", cap, (long long)(intptr_t)p);
}

// managed_alloc(nbytes) -> (capsule, address) of a CUDA managed block: one pointer for NumPy and kernels
static void managed_capsule_free(PyObject* cap) { nte_free_managed(PyCapsule_GetPointer(cap, "pnumpy_b200.managed")); }
static PyObject* managed_alloc(PyObject*, PyObject* args) {
    long long nbytes = 0;
    if (!PyArg_ParseTuple(args, "L:managed_alloc", &nbytes)) return NULL;
    if (nbytes < 0) { PyErr_SetString(PyExc_ValueError, "negative size"); return NULL; }
    void* p = nte_alloc_managed((size_t)nbytes);
    if (!p) return PyErr_Format(PyExc_MemoryError, "cudaMallocManaged(%lld) failed: %s", nbytes, pncu_last_error());
    PyObject* cap = PyCapsule_New(p, "pnumpy_b200.managed", managed_capsule_free);
    if (!cap) { nte_free_managed(p); return NULL; }
    return Py_BuildValue("(NL)", cap, (long long)(intptr_t)p);
}

// =======================================================================================================
// Fused chain (an ADDITION, the reference has nothing like it): a*b + c and
// sum(a*b + c) in one pass over the operands.  The Python wrappers (pnumpy_b200.muladd / muladd_sum)
// run the ordinary hooked ufunc chain whenever these return False / None, and the results are
// bit-identical either way (include/pnumpy_cuda.h: pncu_any_three_fn, pncu_three_reduce_fn).
static bool fused_operands(PyObject* const* objs, int count, int writable_index, PyArrayObject** arr, int* atype, npy_intp* n) {
    for (int k = 0; k < count; ++k) {
        if (!PyArray_Check(objs[k])) return false;
        arr[k] = (PyArrayObject*)objs[k];
        const int need = NPY_ARRAY_C_CONTIGUOUS | NPY_ARRAY_ALIGNED | (k == writable_index ? NPY_ARRAY_WRITEABLE : 0);
        if ((PyArray_FLAGS(arr[k]) & need) != need) return false;
        if (!PyArray_ISNOTSWAPPED(arr[k])) return false;
        if (PyArray_TYPE(arr[k]) != PyArray_TYPE(arr[0])) return false;
        if (PyArray_NDIM(arr[k]) != PyArray_NDIM(arr[0])) return false;
        for (int d = 0; d < PyArray_NDIM(arr[0]); ++d)
            if (PyArray_DIM(arr[k], d) != PyArray_DIM(arr[0], d)) return false;
    }
    const int dt = PyArray_TYPE(arr[0]);
    if (dt < 0 || dt >= COUNT_OF(convert_dtype_to_atop)) return false;
    *atype = convert_dtype_to_atop[dt];
    *n = PyArray_SIZE(arr[0]);
    return *atype >= 0;
}

static PyObject* fused_muladd(PyObject*, PyObject* args) {
    PyObject* o[4];
    if (!PyArg_ParseTuple(args, "OOOO:fused_muladd", &o[0], &o[1], &o[2], &o[3])) return NULL;
    PyArrayObject* a[4];
    int atype = -1;
    npy_intp n = 0;
    if (!g_Settings.AtopEnabled || !fused_operands(o, 4, 3, a, &atype, &n)) Py_RETURN_FALSE;
    int rc;
    Py_BEGIN_ALLOW_THREADS
    rc = nte_muladd(atype, PyArray_DATA(a[0]), PyArray_DATA(a[1]), PyArray_DATA(a[2]), PyArray_DATA(a[3]), (int64_t)n, 3);
    Py_END_ALLOW_THREADS
    if (rc == NTE_DECLINED) Py_RETURN_FALSE;
    if (rc) return PyErr_Format(PyExc_RuntimeError, "pnumpy_b200: fused multiply-add failed on the GPU (status %d): %s", rc, pncu_last_error());
    ledger(CAT_BINARY, true, n);
    Py_RETURN_TRUE;
}

static PyObject* fused_muladd_sum(PyObject*, PyObject* args) {
    PyObject* o[3];
    if (!PyArg_ParseTuple(args, "OOO:fused_muladd_sum", &o[0], &o[1], &o[2])) return NULL;
    PyArrayObject* a[3];
    int atype = -1;
    npy_intp n = 0;
    if (!g_Settings.AtopEnabled || !fused_operands(o, 3, -1, a, &atype, &n)) Py_RETURN_NONE;
    // ndarray.sum() of a 4-byte integer array accumulates (and returns) int64 / uint64; the fused kernel sums
    // in-type like np.add.reduce(dtype=int32) and ReduceMathOpFast do, so those dtypes take the ufunc chain
    if (atype == PNCU_INT32 || atype == PNCU_UINT32) Py_RETURN_NONE;
    alignas(16) char result[16] = {0};
    int rc;
    Py_BEGIN_ALLOW_THREADS
    rc = nte_muladd_sum(atype, PyArray_DATA(a[0]), PyArray_DATA(a[1]), PyArray_DATA(a[2]), result, (int64_t)n, 3);
    Py_END_ALLOW_THREADS
    if (rc == NTE_DECLINED) Py_RETURN_NONE;
    if (rc) return PyErr_Format(PyExc_RuntimeError, "pnumpy_b200: fused multiply-add-sum failed on the GPU (status %d): %s", rc, pncu_last_error());
    ledger(CAT_BINARY, true, n);
    PyArray_Descr* descr = PyArray_DESCR(a[0]);
    return PyArray_Scalar(result, descr, NULL);
}

// =======================================================================================================
// dtype conversion (the first item of the reference's roadmap,
// doc_src/source/roadmap.rst:11-15; an ADDITION: NumPy offers no public hook for its cast loops).
// convert(src, dst) -> True when dst[...] = src.astype(dst.dtype) ran on the GPU; False when the caller must use
// NumPy's astype (dtype without a kernel, non-contiguous or mismatched operands, below the dispatch threshold).
static PyObject* convert(PyObject*, PyObject* args) {
    PyObject *so = NULL, *dobj = NULL;
    if (!PyArg_ParseTuple(args, "OO:convert", &so, &dobj)) return NULL;
    if (!g_Settings.AtopEnabled || !PyArray_Check(so) || !PyArray_Check(dobj)) Py_RETURN_FALSE;
    PyArrayObject* src = (PyArrayObject*)so;
    PyArrayObject* dst = (PyArrayObject*)dobj;
    const int need = NPY_ARRAY_C_CONTIGUOUS | NPY_ARRAY_ALIGNED;
    if ((PyArray_FLAGS(src) & need) != need || (PyArray_FLAGS(dst) & (need | NPY_ARRAY_WRITEABLE)) != (need | NPY_ARRAY_WRITEABLE)) Py_RETURN_FALSE;
    if (!PyArray_ISNOTSWAPPED(src) || !PyArray_ISNOTSWAPPED(dst) || PyArray_SIZE(src) != PyArray_SIZE(dst)) Py_RETURN_FALSE;
    const int st = PyArray_TYPE(src), dt = PyArray_TYPE(dst);
    if (st < 0 || st >= COUNT_OF(convert_dtype_to_atop) || dt < 0 || dt >= COUNT_OF(convert_dtype_to_atop)) Py_RETURN_FALSE;
    const int sa = convert_dtype_to_atop[st], da = convert_dtype_to_atop[dt];
    if (sa < 0 || da < 0) Py_RETURN_FALSE;
    pncu_unary_fn fn = pncu_GetConvertFast(sa, da);
    if (!fn) Py_RETURN_FALSE;
    const npy_intp n = PyArray_SIZE(src);
    const uint64_t t0 = ledger_t0();
    int rc;
    Py_BEGIN_ALLOW_THREADS
    rc = nte_unary_class(fn, pncu_itemsize(sa), pncu_itemsize(da), PyArray_DATA(src), PyArray_DATA(dst), (int64_t)n,
                         pncu_itemsize(sa), pncu_itemsize(da), 3, NTE_CLASS_STREAM);
    Py_END_ALLOW_THREADS
    if (rc == NTE_DECLINED) Py_RETURN_FALSE;
    if (rc) return PyErr_Format(PyExc_RuntimeError, "pnumpy_b200: dtype conversion failed on the GPU (status %d): %s", rc, pncu_last_error());
    ledger(CAT_UNARY, true, n, t0, 0, sa, (int64_t)n * (pncu_itemsize(sa) + pncu_itemsize(da)));
    Py_RETURN_TRUE;
}

// pointer_kind(address) -> 0 pageable host, 1 pinned host, 2 device, 3 managed
static PyObject* pointer_kind(PyObject*, PyObject* args) {
    unsigned long long addr = 0;
    if (!PyArg_ParseTuple(args, "K:pointer_kind", &addr)) return NULL;
    int dev = -1;
    return PyLong_FromLong(pncu_pointer_kind((const void*)(uintptr_t)addr, &dev));
}

static PyMethodDef module_functions[] = {
    {"pointer_kind", pointer_kind, METH_VARARGS, "0 pageable host, 1 pinned host, 2 device, 3 managed"},
    {"convert", convert, METH_VARARGS, "dst[...] = src.astype(dst.dtype) on the GPU; False when the caller must use NumPy's astype"},
    {"fused_muladd", fused_muladd, METH_VARARGS, "a*b+c into out in one GPU pass; False when the caller must run the ufunc chain"},
    {"fused_muladd_sum", fused_muladd_sum, METH_VARARGS, "sum(a*b+c) in one GPU pass; None when the caller must run the ufunc chain"},
    {"managed_alloc", managed_alloc, METH_VARARGS, "(capsule, address) of a CUDA managed block"},
    {"initialize", (PyCFunction)(void (*)(void))newinit, METH_VARARGS | METH_KEYWORDS,
     "Initialize the module. Replaces the ufunc inner loops with wrapped versions using "
     "PyUFunc_ReplaceLoopBySignature and calls numpy.setbufsize(8192 * 1024)."},
    {"atop_enable", atop_enable, METH_VARARGS, "enable the atop (GPU) inner loop implementations."},
    {"atop_disable", atop_disable, METH_VARARGS, "disable the atop (GPU) inner loop implementations."},
    {"atop_isenabled", atop_isenabled, METH_VARARGS, "returns True if atop is enabled"},
    {"atop_info", atop_info, METH_VARARGS, "return dict"},
    {"atop_setworkers", atop_setworkers, METH_VARARGS, "set workers for a func"},
    {"thread_enable", thread_enable, METH_VARARGS, "enable sharding a call over several lanes / GPUs"},
    {"thread_disable", thread_disable, METH_VARARGS, "run every call on one lane of GPU 0"},
    {"thread_isenabled", thread_isenabled, METH_VARARGS, "returns True if threading is enabled"},
    {"thread_getworkers", thread_getworkers, METH_VARARGS, "number of extra lanes a call may use"},
    {"thread_setworkers", thread_setworkers, METH_VARARGS, "set the number of extra lanes; returns the previous value"},
    {"thread_zigzag", thread_zigzag, METH_VARARGS, "toggle zigzag mode"},
    {"timer_gettsc", timer_gettsc, METH_VARARGS, "time stamp counter"},
    {"timer_getutc", timer_getutc, METH_VARARGS, "nanoseconds since the unix epoch"},
    {"ledger_enable", ledger_enable, METH_VARARGS, "enable the ledger"},
    {"ledger_disable", ledger_disable, METH_VARARGS, "disable the ledger"},
    {"ledger_isenabled", ledger_isenabled, METH_VARARGS, "returns True if the ledger is enabled"},
    {"ledger_info", ledger_info, METH_VARARGS, "per-category call counts recorded while the ledger was enabled"},
    {"ledger_records", ledger_records, METH_VARARGS, "the most recent per-call records (op, type, elements, bytes, ticks, gpu)"},
    {"ledger_clear", ledger_clear, METH_VARARGS, "zero the ledger counters and drop the records"},
    {"recycler_enable", (PyCFunction)(void (*)(void))recycler_enable, METH_VARARGS | METH_KEYWORDS,
     "recycler_enable(kind='pinned'|'managed', threshold=None): large ndarrays are born in recycled pinned host / CUDA managed memory"},
    {"recycler_trim", recycler_trim, METH_VARARGS, "free the pooled (unused) blocks; returns how many"},
    {"cuda_trace", cuda_trace, METH_VARARGS, "cuda_trace(on=None): toggle the per-call trace; returns the last resident call's eight stations in ns"},
    {"cuda_inject_fault", cuda_inject_fault, METH_VARARGS, "fail the n-th GPU dispatch from now (tests); returns the previous countdown"},
    {"recycler_disable", recycler_disable, METH_VARARGS, "restore NumPy's default data allocator"},
    {"recycler_isenabled", recycler_isenabled, METH_VARARGS, "returns True if the recycler is enabled"},
    {"recycler_info", recycler_info, METH_VARARGS, "pinned-block pool statistics"},
    {"cpustring", cpustring, METH_VARARGS, "device identification string"},
    {"cuda_info", cuda_info, METH_VARARGS, "devices, thresholds and dispatch counters"},
    {"cuda_reset_stats", cuda_reset_stats, METH_VARARGS, "zero the dispatch counters"},
    {"cuda_setthreshold", cuda_setthreshold, METH_VARARGS, "minimum elements for a GPU dispatch; returns the previous value"},
    {"cuda_getthreshold", cuda_getthreshold, METH_VARARGS, "minimum elements for a GPU dispatch"},
    {"cuda_setslab", cuda_setslab, METH_VARARGS, "bytes per staging slab; returns the previous value"},
    {"pinned_alloc", pinned_alloc, METH_VARARGS, "(capsule, address) of a pinned host block"},
    {NULL, NULL, 0, NULL}};

static PyModuleDef moduledef = {PyModuleDef_HEAD_INIT, "pnumpy_b200._pnumpy",
                                "B200 loops for NumPy ufuncs behind pnumpy's hooks", -1, module_functions,
                                NULL, NULL, NULL, NULL};

PyMODINIT_FUNC PyInit__pnumpy(void) {
    PyObject* module = PyModule_Create(&moduledef);
    if (!module) return NULL;
    import_array();
    import_umath();
    return module;
}
