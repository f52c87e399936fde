"""CPU tests of the drop-in boundary: libpnumpy_cuda.so loads without a GPU, exports every symbol
include/pnumpy_cuda.h declares, hands out the reference's hook table, and fails LOUDLY (no CPU
fallback) when asked to compute without a device."""
import ctypes
import json
import os
import re

import numpy as np
import pytest

from pnumpy_b200 import cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _no_gpu():
    n = ctypes.c_int(0)
    return cabi.load().pncu_init(ctypes.byref(n)) != 0


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "pnumpy_cuda.h")).read()
    declared = set(re.findall(r"PNCU_API[^;]*?\b(pncu_\w+)\s*\(", hdr))
    assert len(declared) >= 40
    lib = ctypes.CDLL(cabi.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    # the ctypes binding covers the same set
    assert declared == set(cabi.SIGNATURES), declared ^ set(cabi.SIGNATURES)


def test_enum_values_match_reference():
    """Enum VALUES are baked into the stub tables (src/atop/common_inc.h:236-372)."""
    hdr = open(os.path.join(ROOT, "include", "pnumpy_cuda.h")).read()
    vals = {k: int(v) for k, v in re.findall(r"\b(PNCU_[A-Z0-9_]+)\s*=\s*(-?\d+)", hdr)}
    want = dict(PNCU_BOOL=0, PNCU_INT64=7, PNCU_UINT64=8, PNCU_FLOAT=12, PNCU_DOUBLE=13,
                PNCU_LONGDOUBLE=14, PNCU_VOID=21, PNCU_ATYPE_LAST=22,
                PNCU_CMP_EQ=0, PNCU_CMP_NE=1, PNCU_CMP_LT=2, PNCU_CMP_GT=3, PNCU_CMP_LTE=4, PNCU_CMP_GTE=5,
                PNCU_ABS=1, PNCU_SQRT=15, PNCU_ISNAN=20, PNCU_BITWISE_NOT=28, PNCU_UNARY_LAST=35,
                PNCU_ADD=1, PNCU_SUB=2, PNCU_MUL=3, PNCU_MIN=5, PNCU_MAX=6, PNCU_DIV=13,
                PNCU_BITWISE_AND=21, PNCU_BITWISE_XOR=22, PNCU_BITWISE_OR=23, PNCU_BINARY_LAST=29,
                PNCU_SIN=1, PNCU_COS=2, PNCU_LOG=13, PNCU_EXP=16, PNCU_CBRT=20, PNCU_TRIG_LAST=21)
    for k, v in want.items():
        assert vals[k] == v, k
    assert cabi.BINARY_OPS["ADD"] == 1 and cabi.TRIG_OPS["EXP"] == 16 and cabi.UNARY_OPS["ISNAN"] == 20


def test_getter_table_matches_reference_hook_table():
    """Same out-types as the reference's getters for every (op, type) => the hook layer registers
    the same NumPy loops (atop_info parity, 359 entries); GPU bodies are a superset of the
    reference's fast pointers."""
    table = json.load(open(os.path.join(ROOT, "tests", "golden", "getter_table.json")))
    lib = cabi.load()
    types = list(range(0, 9)) + [12, 13, 14]
    for kind, getter, two in (("binary", lib.pncu_GetSimpleMathOpFast, True),
                              ("compare", lib.pncu_GetComparisonOpFast, True),
                              ("unary", lib.pncu_GetUnaryOpFast, False),
                              ("trig", lib.pncu_GetTrigOpFast, False)):
        for op, row in table[kind].items():
            for t in types:
                out = ctypes.c_int(-1)
                p = getter(int(op), t, t, ctypes.byref(out)) if two else getter(int(op), t, ctypes.byref(out))
                ref_out, ref_fast = row[str(t)]
                assert out.value == ref_out, (kind, op, t, out.value, ref_out)
                if ref_fast and t != 14:  # no long double on the GPU
                    assert p, f"reference has a fast body for {kind} op {op} type {t}, we do not"
    for op, row in table["reduce"].items():
        for t in types:
            if row[str(t)]:
                assert lib.pncu_GetReduceMathOpFast(int(op), t), (op, t)
    # config coverage: every cell of the BASELINE.json sweeps has a GPU body
    for t in (5, 7, 12, 13):
        o = ctypes.c_int(-1)
        assert lib.pncu_GetSimpleMathOpFast(1, t, t, ctypes.byref(o))
        assert lib.pncu_GetSimpleMathOpFast(3, t, t, ctypes.byref(o))
        assert lib.pncu_GetComparisonOpFast(3, t, t, ctypes.byref(o))
        assert lib.pncu_GetSimpleMathOpFast(13, t, t, ctypes.byref(o)) or lib.pncu_GetIntTrueDivide(t)
        for red in (1, 5, 6):
            assert lib.pncu_GetReduceMathOpFast(red, t)
        assert lib.pncu_GetArgMinMaxOpFast(0, t) and lib.pncu_GetArgMinMaxOpFast(1, t)
    for t in (12, 13):
        o = ctypes.c_int(-1)
        for op in (1, 13, 16):
            assert lib.pncu_GetTrigOpFast(op, t, ctypes.byref(o))
        assert lib.pncu_GetUnaryOpFast(15, t, ctypes.byref(o))


def test_itemsize_and_reduce_geometry():
    lib = cabi.load()
    assert [lib.pncu_itemsize(t) for t in (0, 1, 4, 5, 7, 12, 13, 14)] == [1, 1, 2, 4, 8, 4, 8, 0]
    assert lib.pncu_reduce_block_elems() == 65536
    assert lib.pncu_reduce_partial_count(12, 65536) == 4 and lib.pncu_reduce_partial_count(13, 65537) == 9
    assert lib.pncu_reduce_partial_count(1, 65536) == 1 and lib.pncu_reduce_partial_count(12, 0) == 0
    assert lib.pncu_reduce_partial_bytes(1, 12) == 8   # float32 sum accumulates in float64
    assert lib.pncu_reduce_partial_bytes(5, 12) == 4
    assert lib.pncu_reduce_partial_bytes(1, 14) == 0


def test_fails_loudly_without_gpu():
    """No CPU fallback: without a device every compute entry returns an error and says why."""
    if not _no_gpu():
        pytest.skip("a CUDA device is present")
    lib = cabi.load()
    with pytest.raises(cabi.PnumpyCudaError, match="no usable CUDA device"):
        cabi.init()
    fn, out_t = cabi.get_binary(cabi.BINARY_OPS["ADD"], cabi.ATOP_FLOAT)
    assert fn is not None and out_t == cabi.ATOP_FLOAT
    a = np.ones(8, np.float32)
    rc = fn(a.ctypes.data, a.ctypes.data, a.ctypes.data, 8, 4, 4, 4, None)
    assert rc == -1  # PNCU_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.pncu_last_error()
    assert np.all(a == 1)  # untouched
    assert lib.pncu_launch_count() == 0


def test_ctypes_structs_match_the_header(tmp_path):
    """The ctypes mirrors in pnumpy_b200/cabi.py (pncu_peer_table, pncu_fused_exchange) must have the size and the field
    offsets the C compiler gives the structs of include/pnumpy_cuda.h -- a field added on one side only would shift
    every pointer the kernels read."""
    import ctypes
    import subprocess
    from pnumpy_b200 import cabi
    fields = {"pncu_peer_table": (cabi.PeerTable, ["n", "slots", "arrived"]),
              "pncu_fused_exchange": (cabi.FusedExchange, ["n", "self", "all", "root", "slot_base", "total_count", "slots",
                                                           "arrived", "expected", "done_flag", "done_value", "group_offset"])}
    src = ['#include <stdio.h>', '#include <stddef.h>', '#include "pnumpy_cuda.h"', 'int main(void) {']
    for name, (_, fs) in fields.items():
        src.append(f'  printf("{name} %zu", sizeof({name}));')
        for f in fs:
            src.append(f'  printf(" %zu", offsetof({name}, {f}));')
        src.append('  printf("\\n");')
    src += ['  return 0;', '}']
    c = tmp_path / "layout.c"
    c.write_text("\n".join(src))
    exe = str(tmp_path / "layout")
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", exe, str(c)])
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split("\n")
    seen = 0
    for line in out:
        tok = line.split()
        if not tok:
            continue
        cls, fs = fields[tok[0]]
        assert ctypes.sizeof(cls) == int(tok[1]), (tok[0], ctypes.sizeof(cls), tok[1])
        assert [getattr(cls, f).offset for f in fs] == [int(x) for x in tok[2:]], tok[0]
        assert [f for f, _ in cls._fields_] == fs, "a field is missing from the test's list"
        seen += 1
    assert seen == len(fields)
