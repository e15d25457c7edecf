"""The product's printers (pb_format.cpp, incl. the neighbour joining of `tree`) on the CPU: built with a stub for the
one symbol they take from the CUDA library, fed the oracle's result struct (same layout as pb_region_result), compared
byte for byte with the goldens the unmodified reference wrote."""
import ctypes as C
import subprocess

import pytest

import pbtest
from cases import CASES
from popbam_b200 import capi


@pytest.fixture(scope="module")
def fmt():
    so = pbtest.ROOT / "tests" / "_build" / "libfmtharness.so"
    so.parent.mkdir(exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", str(so), str(pbtest.ROOT / "tests" / "fmt_harness.cpp"),
                           str(pbtest.ROOT / "popbam_b200" / "csrc" / "pb_format.cpp")])
    L = C.CDLL(str(so))
    L.fmt_set_params.argtypes = [C.POINTER(capi.Params)]
    L.pb_format_window.restype = C.c_int64
    L.pb_format_window.argtypes = [C.c_void_p, C.POINTER(capi.Result), C.c_int32, C.c_uint32, C.POINTER(capi.PrintOpts), C.c_char_p, C.c_int64]
    return L


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_product_printers_on_oracle_results_equal_reference_text(fmt, case):
    fx, p, an, wb, we, o = pbtest.case_setup(case)
    orc = pbtest.OracleRun(p, fx.batch(), fx.ref(), an, wb, we)
    fmt.fmt_set_params(C.byref(p))
    buf = C.create_string_buffer(1 << 22)
    out = []
    if case[3] == "SNP" and o.snp_output == 2:
        out.append(pbtest.ms_header(fx, len(wb)))
    for w in range(orc.res.n_windows):
        k = fmt.pb_format_window(None, C.byref(orc.res), w, an, C.byref(o), buf, len(buf))
        assert 0 <= k < len(buf)
        out.append(buf.raw[:k].decode())
    ok, why = pbtest.texts_equal("".join(out), pbtest.golden_text(case), snp0=(case[3] == "SNP" and o.snp_output == 0))
    assert ok, why
    orc.close()
