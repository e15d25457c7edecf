"""GPU parity of the counting pileup path (popbam_b200/csrc/pb_pile.cuh, pb_fast.cuh) on shapes chosen to hit its branches:
the same region is run through the counting path, through k_pileup_call (POPBAM_B200_PILEUP=classic) and through
the CPU oracle, and all three must agree bit for bit on the integer results (1e-9 on the fp64 statistics)."""
import os

import numpy as np
import pytest

import pbtest
import popbam_b200
from test_gpu_parity import assert_same, run_gpu

pytestmark = pytest.mark.gpu

AN_NAMES = ["NUCDIV", "SFS", "LD_ZNS", "HAPLO_K", "DIVERGE_POP"]


def _an(names=AN_NAMES):
    an = 0
    for a in names:
        an |= pbtest.AN[a]
    return an


def _ctx(fx, p, an, wb, we, classic, codes=None):
    """One region through a fresh context; the switches are read in pb_create.  codes: POPBAM_B200_CODES (who collects
    the hard cells' base codes: the block's own k_pile_reads CTA, k_cell_codes, or every other block each)."""
    want = {"POPBAM_B200_PILEUP": "classic" if classic else None, "POPBAM_B200_CODES": codes}
    old = {k: os.environ.pop(k, None) for k in want}
    for k, v in want.items():
        if v is not None:
            os.environ[k] = v
    try:
        return run_gpu(fx, p, an, wb, we)
    finally:
        for k in want:
            os.environ.pop(k, None)
            if old[k] is not None:
                os.environ[k] = old[k]


@pytest.mark.parametrize("name,kw,pkw", [
    # several tiles per warp in k_pile_reads, depths beyond the run that a count alone settles: most cells go through the
    # test of the high-quality count or to k_hard_cells
    ("deep", dict(contig_len=9000, n_ingroup=3, has_outgroup=1, depth=60.0, snp_density=0.02, het_frac=0.3, seed=41), {}),
    # variants everywhere, N / = / X / D cigar operations, lower-case and N reference bytes
    ("dense_edge", dict(contig_len=9000, n_ingroup=7, has_outgroup=1, depth=25.0, snp_density=0.3, het_frac=0.5, edge_mode=1, seed=42), {}),
    # 36 bp reads: short segments (the masked first / last pair of the scatter loop is most of the loop)
    ("short", dict(contig_len=9000, n_ingroup=4, has_outgroup=1, depth=18.0, read_len=36, snp_density=0.03, seed=43), {}),
    # 250 bp reads: a warp's tile holds fewer reads than lanes
    ("long", dict(contig_len=12000, n_ingroup=4, has_outgroup=1, depth=30.0, read_len=250, snp_density=0.03, seed=44), {}),
    # reads with mapQ below min_rmsQ contribute: their cells must leave the easy path (the rms test is per cell)
    ("lowq", dict(contig_len=9000, n_ingroup=5, has_outgroup=1, depth=24.0, snp_density=0.03, seed=45), dict(min_rmsQ=38)),
    # min_depth above the mean passing depth, min_baseQ high: coverage decided by the packed depth compare
    ("min_depth", dict(contig_len=9000, n_ingroup=2, has_outgroup=1, depth=30.0, snp_density=0.03, seed=46),
     dict(min_depth=10, min_baseQ=25)),
    # 64 samples: both halves of the 64-bit site masks, sparse coverage (empty cells, strips without records)
    ("n64_sparse", dict(contig_len=6000, n_ingroup=63, has_outgroup=1, depth=14.0, snp_density=0.05, het_frac=0.2, seed=47),
     dict(min_depth=2)),
    # Illumina-1.3 qualities (-i): the packed quality thresholds of k_pile_reads are offset by 31
    ("illumina", dict(contig_len=9000, n_ingroup=5, has_outgroup=1, depth=24.0, snp_density=0.03, seed=49),
     dict(flags=pbtest.FLAG["ILLUMINA"], min_baseQ=4)),
    # odd read length (padding nibble / byte after every read), two read groups per sample
    ("odd_rg2", dict(contig_len=9000, n_ingroup=4, has_outgroup=1, rg_per_sample=2, depth=12.0, read_len=75, snp_density=0.03, seed=50), {}),
    # base-quality threshold above every quality value but the two highest
    ("baseq", dict(contig_len=9000, n_ingroup=4, has_outgroup=1, depth=40.0, snp_density=0.03, seed=51), dict(min_baseQ=33)),
    # heterozygote mode
    ("het", dict(contig_len=9000, n_ingroup=6, has_outgroup=1, depth=28.0, snp_density=0.05, het_frac=0.6, seed=48),
     dict(flags=pbtest.FLAG["HETEROZYGOTE"])),
])
def test_counting_path_equals_classic_kernel_and_oracle(name, kw, pkw):
    fx = pbtest.Fixture(**kw)
    pkw = dict(pkw)
    flags = pkw.pop("flags", 0)
    p = fx.params(flags=flags | pbtest.FLAG["OUTGROUP"], outidx=fx.n_samples - 1, **pkw)
    wb, we = pbtest.window_grid(0, fx.contig_len, 3000)
    an = 0
    for a in AN_NAMES:
        an |= pbtest.AN[a]
    fast = _ctx(fx, p, an, wb, we, classic=False)
    classic = _ctx(fx, p, an, wb, we, classic=True)
    assert fast.path() == 1 and classic.path() == 0, "counting path was not taken"
    orc = pbtest.OracleRun(p, fx.batch(), fx.ref(), an, wb, we)
    got, ref, want = pbtest.result_arrays(fast.res), pbtest.result_arrays(classic.res), pbtest.result_arrays(orc.res)
    assert int(want["segsites"].sum()) > 0
    assert_same(got, ref, AN_NAMES)
    assert_same(got, want, AN_NAMES)
    orc.close(); fast.close(); classic.close(); fx.close()


@pytest.mark.parametrize("name,pkw,windows", [
    # no read passes min_mapQ: no read is counted, every cell is empty
    ("no_usable_reads", dict(min_mapQ=255), [(0, 3000), (3000, 6000)]),
    # windows with gaps between them, starting at an odd offset: position blocks straddle window edges and gap positions
    ("gaps_unaligned", {}, [(1237, 2001), (2500, 4999), (6001, 8000)]),
    # a region shorter than one strip
    ("tiny", {}, [(4111, 4130)]),
    # one-position windows
    ("single_positions", {}, [(100, 101), (5000, 5001), (8999, 9000)]),
])
def test_counting_path_region_shapes(name, pkw, windows):
    fx = pbtest.Fixture(contig_len=9000, n_ingroup=5, has_outgroup=1, depth=22.0, snp_density=0.05, het_frac=0.3, seed=61)
    p = fx.params(flags=pbtest.FLAG["OUTGROUP"], outidx=fx.n_samples - 1, **pkw)
    wb = np.array([w[0] for w in windows], dtype=np.int32); we = np.array([w[1] for w in windows], dtype=np.int32)
    an = 0
    for a in AN_NAMES:
        an |= pbtest.AN[a]
    fast = _ctx(fx, p, an, wb, we, classic=False)
    classic = _ctx(fx, p, an, wb, we, classic=True)
    assert fast.path() == 1 and classic.path() == 0, "counting path was not taken"
    orc = pbtest.OracleRun(p, fx.batch(), fx.ref(), an, wb, we)
    got, ref, want = pbtest.result_arrays(fast.res), pbtest.result_arrays(classic.res), pbtest.result_arrays(orc.res)
    assert_same(got, ref, AN_NAMES)
    assert_same(got, want, AN_NAMES)
    orc.close(); fast.close(); classic.close(); fx.close()


@pytest.mark.parametrize("name,kw", [
    ("dense_edge", dict(contig_len=9000, n_ingroup=7, has_outgroup=1, depth=25.0, snp_density=0.3, het_frac=0.5, edge_mode=1, seed=42)),
    ("n64_sparse", dict(contig_len=6000, n_ingroup=63, has_outgroup=1, depth=14.0, snp_density=0.05, het_frac=0.2, seed=47)),
    # two samples, deep: long per-sample read lists in a block
    ("deep_few", dict(contig_len=9000, n_ingroup=1, has_outgroup=1, depth=60.0, snp_density=0.1, het_frac=0.5, seed=52)),
])
def test_hard_cell_codes_collected_in_the_block_or_by_k_cell_codes(name, kw):
    """The base codes of the cells left for k_hard_cells are collected by the block's own k_pile_reads CTA (reads fresh in
    L2) or, when its lists lack the room, by k_cell_codes: all in the block, all by k_cell_codes, and every other block each
    must give the same results, equal to the oracle's."""
    fx = pbtest.Fixture(**kw)
    p = fx.params(flags=pbtest.FLAG["OUTGROUP"], outidx=fx.n_samples - 1, min_depth=2)
    wb, we = pbtest.window_grid(0, fx.contig_len, 3000)
    an = _an()
    orc = pbtest.OracleRun(p, fx.batch(), fx.ref(), an, wb, we)
    want = pbtest.result_arrays(orc.res)
    assert int(want["segsites"].sum()) > 0
    for codes in (None, "separate", "mixed"):
        ctx = _ctx(fx, p, an, wb, we, classic=False, codes=codes)
        assert ctx.path() == 1, "counting path was not taken"
        assert_same(pbtest.result_arrays(ctx.res), want, AN_NAMES)
        ctx.close()
    orc.close(); fx.close()


def test_depth_cap_binds_and_the_region_takes_the_single_kernel_pileup():
    """The counting path needs "the raw-depth cap can never bind" (k_depth_bound); when it can, the region goes through
    k_pileup_call (sample partition, base codes and quality levels built then) and must still equal the oracle."""
    fx = pbtest.Fixture(contig_len=9000, n_ingroup=5, has_outgroup=1, depth=26.0, snp_density=0.04, het_frac=0.3, seed=71)
    p = fx.params(flags=pbtest.FLAG["OUTGROUP"], outidx=fx.n_samples - 1, max_depth=14)
    wb, we = pbtest.window_grid(0, fx.contig_len, 3000)
    an = 0
    for a in AN_NAMES:
        an |= pbtest.AN[a]
    ctx = run_gpu(fx, p, an, wb, we)
    orc = pbtest.OracleRun(p, fx.batch(), fx.ref(), an, wb, we)
    got, want = pbtest.result_arrays(ctx.res), pbtest.result_arrays(orc.res)
    assert int(want["segsites"].sum()) > 0
    assert_same(got, want, AN_NAMES)
    orc.close(); ctx.close(); fx.close()


NOSYNC = ["NUCDIV", "SFS", "HAPLO_K", "DIVERGE_POP"]     # analyses whose buffers do not depend on the number of segregating sites


def test_asynchronous_regions_equal_the_oracle_and_the_synchronous_run():
    """After its first region a context enqueues the next ones without a host round trip, on launch parameters learnt
    from the earlier regions and verified on the device (pb_lib.cu run_pipeline).  Results must not depend on the mode."""
    fxs = [pbtest.Fixture(contig_len=9000, n_ingroup=5, has_outgroup=1, depth=d, snp_density=0.04, het_frac=0.3, seed=sd)
           for d, sd in ((22.0, 81), (31.0, 82), (12.0, 83))]
    p = fxs[0].params(flags=pbtest.FLAG["OUTGROUP"], outidx=fxs[0].n_samples - 1)
    an = _an(NOSYNC)
    wb, we = pbtest.window_grid(0, 9000, 3000)
    for sync in (False, True):
        if sync:
            os.environ["POPBAM_B200_SYNC"] = "1"
        try:
            ctx = popbam_b200.Context(p)
        finally:
            os.environ.pop("POPBAM_B200_SYNC", None)
        for fx in fxs:
            ctx.set_contig(0, fx.ref())
            ctx.region_begin(an, wb, we)
            ctx.push_batch_async(fx.batch())
            ctx.region_end()
            assert ctx.path() == 1
            orc = pbtest.OracleRun(p, fx.batch(), fx.ref(), an, wb, we)
            assert_same(pbtest.result_arrays(ctx.res), pbtest.result_arrays(orc.res), NOSYNC)
            orc.close()
        # the same reads again, resident on the device
        ctx.relaunch(); ctx.wait()
        orc = pbtest.OracleRun(p, fxs[-1].batch(), fxs[-1].ref(), an, wb, we)
        assert_same(pbtest.result_arrays(ctx.res), pbtest.result_arrays(orc.res), NOSYNC)
        orc.close()
        assert ctx.reruns() == 0
        ctx.close()
    for fx in fxs:
        fx.close()


def test_asynchronous_region_whose_assumptions_fail_is_run_again():
    """Region 2 has longer reads than the context has seen (a warp's tile then holds fewer of them: no failure) and base
    qualities far above the ceiling of the one-stray-base rule (such cells go to k_hard_cells: no failure either);
    region 3 is deep enough for the raw-depth cap to bind.  The device reports that, the host runs the region again
    (with the round trips), and the results still equal the oracle."""
    a = pbtest.Fixture(contig_len=9000, n_ingroup=4, has_outgroup=1, depth=20.0, snp_density=0.04, seed=91)
    b = pbtest.Fixture(contig_len=9000, n_ingroup=4, has_outgroup=1, depth=20.0, read_len=150, snp_density=0.2, het_frac=0.4, edge_mode=1, seed=92)
    c = pbtest.Fixture(contig_len=9000, n_ingroup=4, has_outgroup=1, depth=90.0, snp_density=0.04, seed=93)
    p = a.params(flags=pbtest.FLAG["OUTGROUP"], outidx=a.n_samples - 1, max_depth=110)
    an = _an(NOSYNC)
    wb, we = pbtest.window_grid(0, 9000, 3000)
    ctx = popbam_b200.Context(p)
    paths = []
    for fx in (a, b, c, a):
        ctx.set_contig(0, fx.ref())
        ctx.region_begin(an, wb, we)
        ctx.push_batch(fx.batch())
        ctx.region_end()
        paths.append(ctx.path())
        orc = pbtest.OracleRun(p, fx.batch(), fx.ref(), an, wb, we)
        assert_same(pbtest.result_arrays(ctx.res), pbtest.result_arrays(orc.res), NOSYNC)
        orc.close()
    assert paths == [1, 1, 0, 1], paths          # region 3 ends up in the single-kernel pileup (cap binds)
    assert ctx.reruns() >= 1
    # unsorted reads in asynchronous mode are still reported (bam_pileup.c:384-395)
    import ctypes as C
    from test_gpu_parity import _slice_batch
    keep = []
    bb = a.batch()
    hi = _slice_batch(bb, bb.n_reads // 2, bb.n_reads, keep)
    lo = _slice_batch(bb, 0, bb.n_reads // 2, keep)
    ctx.set_contig(0, a.ref())
    ctx.region_begin(an, wb, we)
    ctx.push_batch(hi); ctx.push_batch(lo)
    assert ctx.L.pb_region_end(ctx.h, C.byref(ctx.res)) == -5
    ctx.close()
    for fx in (a, b, c):
        fx.close()


def test_more_than_8000_live_reads_are_refused():
    """bam_plp_push drops a read once 8000 are live at its start position (bam_pileup.c:260,375); the library does not
    reproduce that and must say so instead of diverging silently: k_depth_bound bounds the live reads from the sorted
    start positions, the region ends with PB_ERR_UNSUPPORTED."""
    import ctypes as C
    fx = pbtest.Fixture(contig_len=2000, n_ingroup=1, has_outgroup=1, depth=4500.0, snp_density=0.01, seed=77)
    p = fx.params(flags=pbtest.FLAG["OUTGROUP"], outidx=fx.n_samples - 1)
    wb, we = pbtest.window_grid(0, fx.contig_len, 1000)
    for sync in (True, False):
        if sync:
            os.environ["POPBAM_B200_SYNC"] = "1"
        try:
            ctx = popbam_b200.Context(p)
        finally:
            os.environ.pop("POPBAM_B200_SYNC", None)
        ctx.set_contig(0, fx.ref())
        ctx.region_begin(_an(NOSYNC), wb, we)
        ctx.push_batch(fx.batch())
        assert ctx.L.pb_region_end(ctx.h, C.byref(ctx.res)) == -6
        assert b"8000" in ctx.L.pb_last_error(ctx.h)
        ctx.close()
    fx.close()
