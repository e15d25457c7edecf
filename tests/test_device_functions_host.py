"""The product's inlinable device functions (pb_cell.cuh, pb_walk.cuh), compiled for the host by tests/hd_harness.cpp,
against the reference's known-answer vectors.  Checks the histogram formulation of errmod_cal (no sort) and the
per-site logic without a GPU; the same vectors are run through the real kernels in test_gpu_parity.py."""
import ctypes as C
import subprocess

import numpy as np
import pytest

import pbtest
from test_oracle_pin import CELL, SITE


@pytest.fixture(scope="module")
def hd():
    so = pbtest.ROOT / "tests" / "_build" / "libhdharness.so"
    so.parent.mkdir(exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                           "-o", str(so), str(pbtest.ROOT / "tests" / "hd_harness.cpp")])
    L = C.CDLL(str(so))
    L.hd_call_cell.restype = C.c_uint64
    L.hd_call_cell.argtypes = [C.POINTER(C.c_double)] * 3 + [C.POINTER(C.c_uint16), C.c_int, C.c_int, C.c_int]
    L.hd_site_logic.argtypes = [C.POINTER(C.c_uint64)] + [C.c_int] * 7 + [C.POINTER(C.c_uint64)] * 2
    return L


def test_histogram_walk_equals_reference_errmod(hd):
    rec = np.fromfile(pbtest.GOLDEN / "kat_cells.bin", dtype=CELL)
    t = pbtest.oracle_tables()
    for i, r in enumerate(rec):
        if r["k"] == 0:
            continue
        codes = np.array(r["codes"], dtype=np.uint16)
        for r4 in ((i & 3), 0):     # the base rotation must not change the result
            cb = hd.hd_call_cell(*t.ptrs(), codes.ctypes.data_as(C.POINTER(C.c_uint16)), int(r["k"]), int(r["rmsq"]), r4)
            assert cb == int(r["cb"]), "cell %d k=%d" % (i, r["k"])


def test_site_logic_equals_reference(hd):
    rec = np.fromfile(pbtest.GOLDEN / "kat_sites.bin", dtype=SITE)
    for r in rec:
        n = int(r["n"])
        cb = np.array(r["cb_in"], dtype=np.uint64)
        cov, typ = C.c_uint64(), C.c_uint64()
        fq = hd.hd_site_logic(cb.ctypes.data_as(C.POINTER(C.c_uint64)), n, int(r["ref"]), int(r["het"]), int(r["min_snpq"]),
                              int(r["min_rmsq"]), int(r["min_depth"]), int(r["max_depth"]), C.byref(cov), C.byref(typ))
        assert fq == int(r["fq"])
        assert cov.value == int(r["cov"])
        assert (cb[:n] == r["cb_out"][:n]).all()
