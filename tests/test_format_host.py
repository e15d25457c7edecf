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


def test_neighbour_joining_agrees_with_the_oracle_on_random_matrices(fmt):
    """The product's array-based neighbour joining (pb_format.cpp) against the oracle's ring-of-nodes restatement on random
    difference matrices full of ties (small integers), p-distance and Jukes-Cantor: same topology, same branch lengths."""
    import numpy as np
    L = pbtest.oracle_lib()
    rng = np.random.default_rng(11)
    buf_a, buf_b = C.create_string_buffer(1 << 16), C.create_string_buffer(1 << 16)
    for trial in range(300):
        n = int(rng.integers(2, 24))
        T = n + 1
        hi = int(rng.choice([2, 4, 30, 400]))
        d = rng.integers(0, hi + 1, size=(T, T)).astype(np.uint16)
        d = np.triu(d, 1); d = (d + d.T).astype(np.uint16)
        res = capi.Result()
        res.n_windows, res.n_pops, res.n_samples, res.analyses = 1, 1, n, pbtest.AN["TREE"]
        wb = (C.c_int32 * 1)(0); we = (C.c_int32 * 1)(10000)
        ns = (C.c_int32 * 1)(int(rng.integers(500, 10000))); sg = (C.c_int32 * 1)(int(rng.integers(1, 50)))
        so = (C.c_int64 * 2)(0, 0)
        res.win_beg, res.win_end, res.num_sites, res.segsites, res.seg_off = wb, we, ns, sg, so
        flat = np.ascontiguousarray(d.reshape(-1))
        res.tree_diff = flat.ctypes.data_as(C.POINTER(C.c_uint16))
        names = [("s%d" % i).encode() for i in range(n)]
        o = capi.PrintOpts()
        keep = [(C.c_char_p * 1)(b"pop"), (C.c_char_p * n)(*names)]
        o.chrom, o.pop_names, o.sample_names = b"chr1", keep[0], keep[1]
        o.min_sites, o.min_snps, o.jc, o.snp_output, o.ref_name = 10, 10, trial & 1, 0, b"ref"
        p = capi.Params()
        p.n_samples, p.n_pops = n, 1
        fmt.fmt_set_params(C.byref(p))
        ka = fmt.pb_format_window(None, C.byref(res), 0, pbtest.AN["TREE"], C.byref(o), buf_a, len(buf_a))
        kb = L.pbo_format_window(C.byref(p), C.byref(res), 0, pbtest.AN["TREE"], C.byref(o), buf_b, len(buf_b))
        assert ka == kb and ka > 0
        assert buf_a.raw[:ka] == buf_b.raw[:kb], (trial, n, buf_a.raw[:ka], buf_b.raw[:kb])
