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
    L.hd_by_count.restype = C.c_long
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


def test_unanimous_shortcut_equals_oracle(hd):
    """Cells whose bases all agree take the early-exit walk (pb_walk_unanimous): compare with the reference-pinned
    oracle on random cells over the whole depth range and several quality distributions."""
    L, t = pbtest.oracle_lib(), pbtest.oracle_tables()
    rng = np.random.default_rng(42)
    n_unan = 0
    for i in range(6000):
        style = i % 4
        k = int(rng.integers(1, 256)) if style == 0 else int(rng.integers(1, 40))
        if style == 1:
            q = rng.choice([20, 29, 30, 35, 40], size=k)
        elif style == 2:
            q = rng.integers(4, 64, size=k)
        elif style == 3:
            q = rng.choice([4, 5, 6], size=k)            # weak evidence: the shortcut may not trigger
        else:
            q = rng.choice([13, 25, 37, 41, 60, 63], size=k)
        b = int(rng.integers(0, 4))
        base = np.full(k, b)
        if i % 7 == 0 and k > 1:                          # a few non-unanimous cells through the same entry point
            base[rng.integers(0, k)] = (b + 1) & 3
        strand = rng.integers(0, 2, size=k)
        codes = (q << 5 | strand << 4 | base).astype(np.uint16)
        rmsq = int((rng.integers(13, 61, size=k) ** 2).sum())
        work = np.zeros(256, dtype=np.uint16); work[:k] = codes
        want = L.pbo_call_cell(*t.ptrs(), work.ctypes.data_as(C.POINTER(C.c_uint16)), k, rmsq, None)
        full = np.zeros(256, dtype=np.uint16); full[:k] = codes
        got = hd.hd_call_cell(*t.ptrs(), full.ctypes.data_as(C.POINTER(C.c_uint16)), k, rmsq, int(rng.integers(0, 4)))
        assert got == want, (i, k, b)
        n_unan += len(set(base.tolist())) == 1
    assert n_unan > 4000
    assert hd.hd_by_count() > 2000          # the walk-free count test decided most of them


def test_one_stray_base_rule_equals_oracle(hd):
    """pb_one_stray_entry (the counting pass settles cells with k-1 reference bases and one other base when it says
    so): wherever it answers 1, the reference-pinned oracle must call the cell homozygous for the majority base, for
    every quality / strand arrangement tried.  Two rounds per level set: the stray base's level from all levels, and only
    from those below quality 30 (what the kernel knows about a stray base outside the H plane)."""
    L, t = pbtest.oracle_lib(), pbtest.oracle_tables()
    hd.hd_one_stray.argtypes = [C.POINTER(C.c_double)] * 3 + [C.POINTER(C.c_uint8), C.c_int, C.c_int, C.c_int, C.c_int]
    rng = np.random.default_rng(7)
    n_settled = n_low_only = 0
    for levels in ([20, 29, 30, 35, 40], [13, 25, 37, 41, 60, 63], list(range(4, 64)), [4, 5, 6], [40]):
        qv = np.array(levels, dtype=np.uint8)
        qp = qv.ctypes.data_as(C.POINTER(C.c_uint8))
        n_low = sum(1 for q in levels if q < 30)
        full = [hd.hd_one_stray(*t.ptrs(), qp, len(levels), k, 0, len(levels)) for k in range(64)]
        assert full[0] == 0 and full[1] == 0
        rounds = [(levels, full)]
        if 0 < n_low < len(levels):
            low = [hd.hd_one_stray(*t.ptrs(), qp, len(levels), k, 0, n_low) for k in range(64)]
            assert all(a >= b for a, b in zip(low, full))          # fewer stray levels can only widen the range
            n_low_only += sum(a > b for a, b in zip(low, full))
            rounds.append((levels[:n_low], low))
        for stray_levels, ok in rounds:
            for k in range(2, 64):
                if not ok[k]:
                    continue
                for rep in range(12):
                    b = int(rng.integers(0, 4)); e = (b + 1 + int(rng.integers(0, 3))) & 3
                    # adversarial arrangements: the majority at the lowest level, the stray base at its highest; random ones
                    q = rng.choice(levels, size=k) if rep >= 4 else np.full(k, levels[0])
                    base = np.full(k, b); stray = int(rng.integers(0, k)); base[stray] = e
                    q[stray] = stray_levels[-1] if rep < 8 else rng.choice(stray_levels)
                    strand = rng.integers(0, 2, size=k) if rep % 2 else np.full(k, rep // 2 % 2)
                    codes = np.zeros(256, dtype=np.uint16)
                    codes[:k] = (q.astype(np.uint16) << 5 | strand.astype(np.uint16) << 4 | base.astype(np.uint16))
                    cb = L.pbo_call_cell(*t.ptrs(), codes.ctypes.data_as(C.POINTER(C.c_uint16)), k, 900 * k, None)
                    assert (cb >> 8) & 0xff == (b << 2 | b), (levels, stray_levels, k, rep)
                    n_settled += 1
    assert n_settled > 3000 and n_low_only > 0
