"""Host feeder of the `popbam` command line (popbam_b200/csrc/pb_bamio.cpp): BGZF inflate, BAM record decode, BAI
region slicing and @RG sample tables, checked on the CPU against the generator's in-memory batch of the same fixture.
(`popbam _fetch` is a test hook that stops before any GPU call.)"""
import subprocess

import numpy as np
import pytest

import pbtest
import popbam_b200

EXE = popbam_b200.capi.PKG / "_build" / "popbam"


def _load(path):
    raw = open(path, "rb").read()
    n, nc, nb, tid, beg, end = np.frombuffer(raw[:48], dtype=np.int64)
    o = 48
    out = {}
    for name, cnt, dt in (("pos", n, np.int32), ("meta", n, np.uint32), ("cig_off", n + 1, np.uint32), ("cigar", nc, np.uint32),
                          ("base_off", n + 1, np.uint32), ("seq4", nb // 2, np.uint8), ("qual", nb, np.uint8)):
        sz = int(cnt) * np.dtype(dt).itemsize
        out[name] = np.frombuffer(raw[o:o + sz], dtype=dt)
        o += sz
    assert o == len(raw)
    return out


def _expected(fx, beg, end):
    """Records of the generator's batch that bam_fetch would deliver for [beg, end) (is_overlap, bam_index.c:729)."""
    b = fx.batch()
    pos = np.ctypeslib.as_array(b.pos, (b.n_reads,)); meta = np.ctypeslib.as_array(b.meta, (b.n_reads,))
    co = np.ctypeslib.as_array(b.cig_off, (b.n_reads + 1,)); bo = np.ctypeslib.as_array(b.base_off, (b.n_reads + 1,))
    cig = np.ctypeslib.as_array(b.cigar, (b.n_cigar,)); qual = np.ctypeslib.as_array(b.qual, (b.n_bases,))
    seq = np.ctypeslib.as_array(b.seq4, (b.n_bases // 2,))
    ops, lens = cig & 15, cig >> 4
    span = np.where(np.isin(ops, [0, 2, 3, 7, 8]), lens, 0)
    cs = np.concatenate([[0], np.cumsum(span)])
    rend = pos + (cs[co[1:]] - cs[co[:-1]])
    keep = (rend > beg) & (pos < end)
    return pos, meta, co, bo, cig, qual, seq, np.nonzero(keep)[0]


@pytest.mark.parametrize("fxname,region,beg,end", [("edge", "chr1", 0, 20500), ("edge", "chr1:5001-9000", 5000, 9000),
                                                    ("rg2", "chr1:12,001-20,500", 12000, 20500), ("c1", "chr1:30000-30400", 29999, 30400)])
def test_fetch_region_equals_generator_batch(tmp_path, fxname, region, beg, end):
    popbam_b200.build()
    fx = pbtest.fixture(fxname)
    bam, fa = fx.write_files(tmp_path / fxname)
    out = tmp_path / "batch.bin"
    r = subprocess.run([str(EXE), "_fetch", bam, region, str(out)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr
    # sample / population tables (pop_sample.cpp:15-107, popbam.cpp:145-171)
    lines = r.stdout.splitlines()
    assert lines[0].split()[1:] == fx.sample_names
    masks = fx.params().pop_mask
    assert lines[1].split()[1:] == ["%s=%x" % (nm, masks[i]) for i, nm in enumerate(fx.pop_names)]
    got = _load(out)
    pos, meta, co, bo, cig, qual, seq, idx = _expected(fx, beg, end)
    R = fx.sp.read_len
    assert len(got["pos"]) == len(idx)
    assert np.array_equal(got["pos"], pos[idx])
    assert np.array_equal(got["meta"], meta[idx])          # flag, mapq and the RG -> sample lookup
    assert np.array_equal(np.diff(got["cig_off"].astype(np.int64)), (co[idx + 1] - co[idx]).astype(np.int64))
    assert np.array_equal(got["cigar"], np.concatenate([cig[co[i]:co[i + 1]] for i in idx]) if len(idx) else got["cigar"])
    assert (got["base_off"] % 4 == 0).all()
    for k in list(range(min(50, len(idx)))) + list(range(max(0, len(idx) - 50), len(idx))):
        i = idx[k]
        g0 = int(got["base_off"][k])
        assert np.array_equal(got["qual"][g0:g0 + R], qual[bo[i]:bo[i] + R])
        assert np.array_equal(got["seq4"][g0 // 2:g0 // 2 + (R + 1) // 2], seq[bo[i] // 2:bo[i] // 2 + (R + 1) // 2])


def test_cli_errors_without_gpu(tmp_path):
    popbam_b200.build()
    r = subprocess.run([str(EXE), "nucdiv", "-f", "nope.fa", str(tmp_path / "missing.bam"), "chr1"], stderr=subprocess.PIPE, text=True)
    assert r.returncode == 1 and "does not exist" in r.stderr
    r = subprocess.run([str(EXE), "frobnicate"], stderr=subprocess.PIPE, text=True)
    assert r.returncode == 1 and "unrecognized command" in r.stderr
