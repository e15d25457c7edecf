"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the C ABI, against
  * the reference's own outputs (tests/golden/*.txt, kat_cells.bin), and
  * the reference-pinned CPU oracle on the same seeded inputs (every result array).
Integer / bit results must be identical; floating-point statistics within 1e-9 relative (BASELINE.json)."""
import ctypes as C

import numpy as np
import pytest

import pbtest
import popbam_b200
from cases import CASES
from test_oracle_pin import CELL

pytestmark = pytest.mark.gpu

RTOL = 1e-9
INT_KEYS = ["win_beg", "win_end", "num_sites", "segsites", "seg_off", "seg_pos", "seg_idx", "seg_type", "seg_ref"]
BY_AN = pbtest.BY_AN


def run_gpu(fx, p, an, wb, we, batches=None):
    ctx = popbam_b200.Context(p)
    ctx.set_contig(0, fx.ref())
    ctx.region_begin(an, wb, we)
    for b in (batches or [fx.batch()]):
        ctx.push_batch(b)
    ctx.region_end()
    return ctx


def assert_same(got, want, an_names):
    for k in INT_KEYS:
        assert np.array_equal(got[k], want[k]), k
    for name in an_names:
        ints, flts = BY_AN[name]
        for k in ints:
            assert np.array_equal(got[k], want[k]), k
        for k in flts:
            assert got[k].shape == want[k].shape, k
            assert np.allclose(got[k], want[k], rtol=RTOL, atol=0.0, equal_nan=True), (k, got[k], want[k])


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_case_matches_reference_golden_and_oracle(case):
    fx, p, an, wb, we, o = pbtest.case_setup(case)
    ctx = run_gpu(fx, p, an, wb, we)
    got_text = ctx.text(an, o)
    if an == pbtest.AN["SNP"] and o.snp_output == 2:
        got_text = pbtest.ms_header(fx, len(wb)) + got_text
    ok, why = pbtest.texts_equal(got_text, pbtest.golden_text(case), snp0=(an == pbtest.AN["SNP"] and o.snp_output == 0))
    assert ok, why
    orc = pbtest.OracleRun(p, fx.batch(), fx.ref(), an, wb, we)
    assert_same(pbtest.result_arrays(ctx.res), pbtest.result_arrays(orc.res), [case[3]])
    assert ctx.res.reads_used == orc.res.reads_used and ctx.res.aligned_bases == orc.res.aligned_bases
    orc.close(); ctx.close()


ALL_AN = 0xfff


@pytest.mark.parametrize("fxname", ["c1", "edge", "rg2", "ld", "n64"])
def test_all_analyses_in_one_pass_with_cb_words(fxname):
    """Every analysis from one pileup pass, plus the per-(site,sample) cb words of the whole span (bit-exact)."""
    fx = pbtest.fixture(fxname)
    p = fx.params(flags=pbtest.FLAG["EMIT_CB"] | pbtest.FLAG["OUTGROUP"], outidx=fx.n_samples - 1)
    wb, we = pbtest.window_grid(0, fx.contig_len, 5000)
    ctx = run_gpu(fx, p, ALL_AN, wb, we)
    orc = pbtest.OracleRun(p, fx.batch(), fx.ref(), ALL_AN, wb, we)
    got, want = pbtest.result_arrays(ctx.res, want_cb=True), pbtest.result_arrays(orc.res, want_cb=True)
    # the oracle only visits positions inside a window (make_X's t.beg <= pos < t.end test); the kernel also calls the
    # gap positions between windows but never uses them
    span = int(we[-1] - wb[0])
    inwin = np.zeros(span, dtype=bool)
    for b, e in zip(wb, we):
        inwin[b - wb[0]:e - wb[0]] = True
    n = fx.n_samples
    assert np.array_equal(got["cb"].reshape(span, n)[inwin], want["cb"].reshape(span, n)[inwin])
    assert np.array_equal(got["site_type"][inwin], want["site_type"][inwin])
    assert np.array_equal(got["site_flag"] & 3, want["site_flag"] & 3)
    assert (got["site_flag"][~inwin] == 0).all()
    assert_same(got, want, list(BY_AN))
    orc.close(); ctx.close()


def _slice_batch(b, lo, hi, keep):
    """Sub-batch [lo,hi) of a pb_read_batch as fresh numpy arrays (offsets rebased to the slice)."""
    pos = np.ctypeslib.as_array(b.pos, (b.n_reads,))[lo:hi].copy()
    meta = np.ctypeslib.as_array(b.meta, (b.n_reads,))[lo:hi].copy()
    co = np.ctypeslib.as_array(b.cig_off, (b.n_reads + 1,))[lo:hi + 1].copy()
    bo = np.ctypeslib.as_array(b.base_off, (b.n_reads + 1,))[lo:hi + 1].copy()
    cig = np.ctypeslib.as_array(b.cigar, (b.n_cigar,))[co[0]:co[-1]].copy()
    qual = np.ctypeslib.as_array(b.qual, (b.n_bases,))[bo[0]:bo[-1]].copy()
    seq = np.ctypeslib.as_array(b.seq4, (b.n_bases // 2,))[bo[0] // 2:bo[-1] // 2].copy()
    co -= co[0]; bo -= bo[0]
    nb = pbtest.Batch()
    nb.n_reads, nb.n_cigar, nb.n_bases = hi - lo, len(cig), len(qual)
    for name, a, t in (("pos", pos, C.c_int32), ("meta", meta, C.c_uint32), ("cig_off", co, C.c_uint32), ("cigar", cig, C.c_uint32),
                       ("base_off", bo, C.c_uint32), ("seq4", seq, C.c_uint8), ("qual", qual, C.c_uint8)):
        keep.append(a)
        setattr(nb, name, a.ctypes.data_as(C.POINTER(t)))
    return nb


def test_multiple_pushes_equal_one_push():
    fx = pbtest.fixture("edge")
    p = fx.params()
    wb, we = pbtest.window_grid(0, fx.contig_len, 10000)
    an = pbtest.AN["NUCDIV"] | pbtest.AN["SFS"] | pbtest.AN["SNP"]
    one = run_gpu(fx, p, an, wb, we)
    b = fx.batch()
    keep = []
    cuts = [0, 1, b.n_reads // 3, b.n_reads // 3, (2 * b.n_reads) // 3 + 7, b.n_reads]
    parts = [_slice_batch(b, cuts[i], cuts[i + 1], keep) for i in range(len(cuts) - 1)]
    many = run_gpu(fx, p, an, wb, we, batches=parts)
    assert_same(pbtest.result_arrays(many.res), pbtest.result_arrays(one.res), ["NUCDIV", "SFS", "SNP"])
    one.close(); many.close()


def test_context_reuse_and_relaunch():
    """Two regions through one context (buffers are reused) and a relaunch on resident reads give the same answer."""
    fx = pbtest.fixture("c1")
    p = fx.params()
    an = pbtest.AN["NUCDIV"]
    wb, we = pbtest.window_grid(0, fx.contig_len, 10000)
    ctx = popbam_b200.Context(p)
    ctx.set_contig(0, fx.ref())
    first = None
    for rep in range(2):
        ctx.region_begin(an, wb, we)
        ctx.push_batch(fx.batch())
        ctx.region_end()
        cur = pbtest.result_arrays(ctx.res)
        if first is None:
            first = cur
        assert_same(cur, first, ["NUCDIV"])
    ctx.relaunch()
    ctx.wait()
    assert_same(pbtest.result_arrays(ctx.res), first, ["NUCDIV"])
    ms = ctx.stage_times()
    assert ms[1] > 0.0
    # sub-region: windows 1.. only (halo reads before the span are ignored correctly)
    ctx.region_begin(an, wb[1:], we[1:])
    ctx.push_batch(fx.batch())
    ctx.region_end()
    sub = pbtest.result_arrays(ctx.res)
    assert np.array_equal(sub["num_sites"], first["num_sites"][1:])
    assert np.array_equal(sub["seg_type"], first["seg_type"][first["seg_off"][1]:])
    assert np.allclose(sub["piw"], first["piw"][fx.n_pops:], rtol=RTOL, atol=0)
    ctx.close()


def test_kernel_cells_known_answers():
    """The reference's errmod_cal/gl2cns vectors through the real pileup kernel: cell i is position i of sample 0,
    covered by k one-base reads whose (quality, strand, base) are the vector's codes (mapq 200 -> qq = baseQ)."""
    rec = np.fromfile(pbtest.GOLDEN / "kat_cells.bin", dtype=CELL)
    rec = rec[rec["k"] > 0]
    n_cells = len(rec)
    ks = rec["k"].astype(np.int64)
    N = int(ks.sum())
    pos = np.repeat(np.arange(n_cells, dtype=np.int32), ks)
    codes = np.concatenate([r["codes"][:r["k"]] for r in rec]).astype(np.uint32)
    strand = (codes >> 4) & 1
    meta = ((strand * 16) << 16 | (200 << 8) | 0).astype(np.uint32)
    cig_off = np.arange(N + 1, dtype=np.uint32)
    cigar = np.full(N, (1 << 4) | 0, dtype=np.uint32)
    base_off = (np.arange(N + 1, dtype=np.uint32) * 2)
    qual = np.zeros(2 * N, dtype=np.uint8); qual[0::2] = (codes >> 5).astype(np.uint8)
    seq4 = ((1 << (codes & 3)) << 4).astype(np.uint8)
    b = pbtest.Batch()
    b.n_reads, b.n_cigar, b.n_bases = N, N, 2 * N
    arrs = dict(pos=(pos, C.c_int32), meta=(meta, C.c_uint32), cig_off=(cig_off, C.c_uint32), cigar=(cigar, C.c_uint32),
                base_off=(base_off, C.c_uint32), seq4=(seq4, C.c_uint8), qual=(qual, C.c_uint8))
    for k, (a, t) in arrs.items():
        setattr(b, k, a.ctypes.data_as(C.POINTER(t)))
    p = pbtest.Params()
    p.n_samples, p.n_pops = 1, 1
    p.pop_mask[0], p.pop_nsmpl[0] = 1, 1
    p.min_depth, p.max_depth, p.min_rmsQ, p.min_snpQ, p.min_mapQ, p.min_baseQ = 0, 255, 0, 0, 0, 0
    # het mode: no clean_heterozygotes; min_snpQ = 0 and a reference byte that matches nothing keeps segbase from
    # reverting; so the emitted word differs from the raw call only in its two flag bits
    p.flags = pbtest.FLAG["EMIT_CB"] | pbtest.FLAG["HETEROZYGOTE"]
    ctx = popbam_b200.Context(p)
    ctx.set_contig(0, b"N" * n_cells)
    ctx.region_begin(0, np.array([0], dtype=np.int32), np.array([n_cells], dtype=np.int32))
    ctx.push_batch(b)
    res = ctx.region_end()
    cb = pbtest.arr(res.cb, n_cells)
    mask = np.uint64(0x0000ffffffffff00)          # snpQ | depth | genotype (rms depends on the synthetic mapq)
    assert np.array_equal(cb & mask, rec["cb"] & mask)
    assert np.array_equal(cb >> np.uint64(48), np.full(n_cells, 200, dtype=np.uint64))
    ctx.close()


def test_push_record_shim_equals_batch():
    """bam_fetch_f-shaped entry point: raw BAM records (core + data) pushed one by one."""
    fx = pbtest.Fixture(contig_len=6000, n_ingroup=3, has_outgroup=1, depth=8.0, edge_mode=1, seed=5)
    p = fx.params()
    wb, we = pbtest.window_grid(0, fx.contig_len, 0)
    an = pbtest.AN["NUCDIV"] | pbtest.AN["SNP"]
    one = run_gpu(fx, p, an, wb, we)
    b = fx.batch()
    ctx = popbam_b200.Context(p)
    ctx.set_contig(0, fx.ref())
    ctx.region_begin(an, wb, we)
    pos = np.ctypeslib.as_array(b.pos, (b.n_reads,)); meta = np.ctypeslib.as_array(b.meta, (b.n_reads,))
    co = np.ctypeslib.as_array(b.cig_off, (b.n_reads + 1,)); bo = np.ctypeslib.as_array(b.base_off, (b.n_reads + 1,))
    cig = np.ctypeslib.as_array(b.cigar, (b.n_cigar,)); qual = np.ctypeslib.as_array(b.qual, (b.n_bases,))
    seq = np.ctypeslib.as_array(b.seq4, (b.n_bases // 2,))
    R = fx.sp.read_len
    for i in range(b.n_reads):
        ncig = int(co[i + 1] - co[i])
        core = np.array([0, pos[i], (int(meta[i] >> 8) & 0xff) << 8 | 3, (int(meta[i]) >> 16) << 16 | ncig, R, -1 & 0xffffffff,
                         -1 & 0xffffffff, 0], dtype=np.uint32)
        data = b"r1\0" + cig[co[i]:co[i + 1]].tobytes() + seq[bo[i] // 2:bo[i] // 2 + (R + 1) // 2].tobytes() + qual[bo[i]:bo[i] + R].tobytes()
        rc = ctx.L.pb_push_record(ctx.h, core.ctypes.data_as(C.c_void_p), data, len(data), int(meta[i] & 0xff))
        assert rc == 0
    ctx.region_end()
    assert_same(pbtest.result_arrays(ctx.res), pbtest.result_arrays(one.res), ["NUCDIV", "SNP"])
    one.close(); ctx.close()


def test_error_paths():
    fx = pbtest.fixture("rg2")
    p = fx.params()
    ctx = popbam_b200.Context(p)
    with pytest.raises(popbam_b200.capi.PopbamError):        # region before contig
        ctx.region_begin(1, [0], [100])
    ctx.set_contig(0, fx.ref())
    with pytest.raises(popbam_b200.capi.PopbamError):        # overlapping windows
        ctx.region_begin(1, [0, 50], [100, 200])
    with pytest.raises(popbam_b200.capi.PopbamError):        # push outside a region
        ctx.push_batch(fx.batch())
    # unsorted input is reported, not silently processed (bam_pileup.c:384-395)
    b = fx.batch()
    keep = []
    hi = _slice_batch(b, b.n_reads // 2, b.n_reads, keep)
    lo = _slice_batch(b, 0, b.n_reads // 2, keep)
    ctx.region_begin(1, [0], [fx.contig_len])
    ctx.push_batch(hi); ctx.push_batch(lo)
    rc = ctx.L.pb_region_end(ctx.h, C.byref(ctx.res))
    assert rc == -5
    # the context stays usable
    ctx.region_begin(1, [0], [fx.contig_len])
    ctx.push_batch(b)
    ctx.region_end()
    assert ctx.res.num_sites[0] > 0
    # empty region: no reads
    ctx.region_begin(0x7ff & ~0x400, [0, 5000], [4000, 9000])
    ctx.region_end()
    assert ctx.res.num_sites[0] == 0 and ctx.res.seg_off[2] == 0
    ctx.close()


@pytest.mark.parametrize("name,kw,pkw", [
    # ~90 segment records per warp and sample: several staging chunks per sample (PB_WCAP), no cap
    ("deep", dict(contig_len=6000, n_ingroup=2, has_outgroup=1, depth=70.0, snp_density=0.02, het_frac=0.3, seed=31), {}),
    # the same pileups with the raw-depth cap binding almost everywhere (CAP kernel variant across chunks)
    ("deep_cap", dict(contig_len=6000, n_ingroup=2, has_outgroup=1, depth=70.0, snp_density=0.02, het_frac=0.3, seed=31),
     dict(max_depth=40)),
    # a variant at every third position, half of them heterozygous: most cells are not unanimous, so the per-warp
    # queues of deferred cells fill up and drain many times
    ("dense", dict(contig_len=8000, n_ingroup=7, has_outgroup=1, depth=25.0, snp_density=0.35, het_frac=0.5, edge_mode=1, seed=32), {}),
    # short reads (36 bp): several reads per 64-byte chunk of the encode pass, many records per position
    ("short", dict(contig_len=8000, n_ingroup=4, has_outgroup=1, depth=30.0, read_len=36, snp_density=0.03, seed=33), {}),
])
def test_stress_shapes_match_oracle(name, kw, pkw):
    fx = pbtest.Fixture(**kw)
    p = fx.params(flags=pbtest.FLAG["EMIT_CB"], **pkw)
    wb, we = pbtest.window_grid(0, fx.contig_len, 2000)
    an = pbtest.AN["NUCDIV"] | pbtest.AN["SFS"] | pbtest.AN["SNP"] | pbtest.AN["LD_ZNS"]
    ctx = run_gpu(fx, p, an, wb, we)
    orc = pbtest.OracleRun(p, fx.batch(), fx.ref(), an, wb, we)
    got, want = pbtest.result_arrays(ctx.res, want_cb=True), pbtest.result_arrays(orc.res, want_cb=True)
    span = int(we[-1] - wb[0])
    inwin = np.zeros(span, dtype=bool)
    for b, e in zip(wb, we):
        inwin[b - wb[0]:e - wb[0]] = True
    n = fx.n_samples
    assert np.array_equal(got["cb"].reshape(span, n)[inwin], want["cb"].reshape(span, n)[inwin])
    assert_same(got, want, ["NUCDIV", "SFS", "SNP", "LD_ZNS"])
    # and without the cb words (the fast paths that skip folding are only taken then)
    p2 = fx.params(**pkw)
    ctx2 = run_gpu(fx, p2, an, wb, we)
    assert_same(pbtest.result_arrays(ctx2.res), want, ["NUCDIV", "SFS", "SNP", "LD_ZNS"])
    orc.close(); ctx.close(); ctx2.close(); fx.close()
