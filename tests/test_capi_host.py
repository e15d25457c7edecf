"""CPU-side checks of the C-ABI library: it loads, exports every symbol of include/popbam_b200.h, its host helpers
(window grid, error-model tables) equal the reference-pinned oracle, and it refuses to run without a CUDA device."""
import ctypes as C
import hashlib
import json
import re

import numpy as np
import pytest

import pbtest
import popbam_b200
from popbam_b200 import capi


@pytest.fixture(scope="module")
def L():
    popbam_b200.build()
    return popbam_b200.lib()


def test_exports_every_declared_symbol(L):
    header = (pbtest.ROOT / "include" / "popbam_b200.h").read_text()
    declared = set(re.findall(r"\b(pb_[a-z_0-9]+)\s*\(", header)) - {"pb_status"}
    assert declared == set(capi.EXPORTS)
    for name in declared:
        assert hasattr(L, name), name


def test_struct_layouts_match_header_sizes():
    # pb_params: 2 ints + 64 u64 + 64 u8 + 10 ints ; the oracle's mirror struct has the same layout
    assert C.sizeof(capi.Params) == 8 + 512 + 64 + 40
    assert C.sizeof(capi.Batch) == 80
    assert C.sizeof(capi.Result) == 16 + 8 * 28 + 8 + 8 * 3 + 8 * 3 + 8          # + tree_diff
    assert C.sizeof(capi.PrintOpts) == 24 + 16 + 8                               # + ref_name


@pytest.mark.parametrize("beg,end,win", [(0, 1000000, 10000), (0, 50000, 10000), (0, 30500, 0), (12345, 99999, 5000),
                                          (0, 10000, 10000), (0, 10001, 10000), (500, 700, 1000)])
def test_window_grid_equals_oracle(L, beg, end, win):
    wb0, we0 = pbtest.window_grid(beg, end, win)
    nw = L.pb_window_grid(beg, end, win, 0, None, None)
    assert nw == len(wb0)
    wb = (C.c_int32 * max(nw, 1))(); we = (C.c_int32 * max(nw, 1))()
    L.pb_window_grid(beg, end, win, nw, wb, we)
    assert list(wb[:nw]) == list(wb0) and list(we[:nw]) == list(we0)


def test_errmod_tables_equal_reference(L):
    fk = np.zeros(256); beta = np.zeros(64 * 65536); lhet = np.zeros(65536)
    dp = C.POINTER(C.c_double)
    assert L.pb_build_errmod_tables(fk.ctypes.data_as(dp), beta.ctypes.data_as(dp), lhet.ctypes.data_as(dp)) == 0
    kat = json.load(open(pbtest.GOLDEN / "kat_tables.json"))
    assert hashlib.sha256(fk.tobytes() + beta.tobytes() + lhet.tobytes()).hexdigest() == kat["sha256"]


def test_errmod_tables_cache_on_disk(L, tmp_path):
    """the tables through the cache -- computed and written the first time, read back the second time,
    rebuilt when the file does not verify; always the reference's tables (sha256 of errmod_init's, kat_tables.json)."""
    dp = C.POINTER(C.c_double)
    kat = json.load(open(pbtest.GOLDEN / "kat_tables.json"))
    d = str(tmp_path / "cache").encode()

    def run():
        fk = np.zeros(256); beta = np.zeros(64 * 65536); lhet = np.zeros(65536)
        hit = C.c_int(-1)
        assert L.pb_errmod_tables_cached_ex(fk.ctypes.data_as(dp), beta.ctypes.data_as(dp), lhet.ctypes.data_as(dp), d, C.byref(hit)) == 0
        assert hashlib.sha256(fk.tobytes() + beta.tobytes() + lhet.tobytes()).hexdigest() == kat["sha256"]
        return hit.value
    assert run() == 0
    files = list((tmp_path / "cache").glob("errmod-v1-*.bin"))
    assert len(files) == 1 and files[0].stat().st_size == 24 + 8 * (256 + 64 * 65536 + 65536)
    assert run() == 1
    raw = bytearray(files[0].read_bytes())
    raw[5000] ^= 0x40                                      # a flipped bit: the checksum does not verify, the tables are rebuilt
    files[0].write_bytes(bytes(raw))
    assert run() == 0
    assert run() == 1


def test_no_cpu_fallback(L):
    """Without a CUDA device pb_create must fail with PB_ERR_CUDA; with one, bad parameters are rejected."""
    import torch
    fx = pbtest.fixture("c1")
    p = fx.params()
    st = C.c_int(0)
    if not torch.cuda.is_available():
        h = L.pb_create(C.byref(p), None, C.byref(st))
        assert not h and st.value == -2
        assert b"no CUDA device" in L.pb_last_error(None)
    p2 = fx.params(max_depth=300)
    h = L.pb_create(C.byref(p2), None, C.byref(st))
    assert not h and st.value == -6
