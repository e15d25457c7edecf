"""GPU parity of the bit-sliced pileup path (popbam_b200/csrc/pb_fast.cuh) on shapes chosen to hit its branches:
the same region is run through the bit-sliced path, through k_pileup_call (POPBAM_B200_PILEUP=classic) and through
the CPU oracle, and all three must agree bit for bit on the integer results (1e-9 on the fp64 statistics)."""
import os

import numpy as np
import pytest

import pbtest
import popbam_b200
from test_gpu_parity import assert_same, run_gpu

pytestmark = pytest.mark.gpu

AN_NAMES = ["NUCDIV", "SFS", "LD_ZNS", "HAPLO_K", "DIVERGE_POP"]


def _ctx(fx, p, an, wb, we, classic):
    old = os.environ.pop("POPBAM_B200_PILEUP", None)
    if classic:
        os.environ["POPBAM_B200_PILEUP"] = "classic"
    try:
        return run_gpu(fx, p, an, wb, we)          # the switch is read in pb_create
    finally:
        os.environ.pop("POPBAM_B200_PILEUP", None)
        if old is not None:
            os.environ["POPBAM_B200_PILEUP"] = old


@pytest.mark.parametrize("name,kw,pkw", [
    # several staging passes per CTA in k_pile_fast (more than 512 records per 64 strips), depths beyond the
    # 6-plane counters' easy range: most cells go through the H-plane test or k_hard_cells
    ("deep", dict(contig_len=9000, n_ingroup=3, has_outgroup=1, depth=60.0, snp_density=0.02, het_frac=0.3, seed=41), {}),
    # variants everywhere, N / = / X / D cigar operations, lower-case and N reference bytes
    ("dense_edge", dict(contig_len=9000, n_ingroup=7, has_outgroup=1, depth=25.0, snp_density=0.3, het_frac=0.5, edge_mode=1, seed=42), {}),
    # 36 bp reads: many records per strip, one staged plane element pair per record
    ("short", dict(contig_len=9000, n_ingroup=4, has_outgroup=1, depth=18.0, read_len=36, snp_density=0.03, seed=43), {}),
    # 250 bp reads: nine plane elements per record, fewer records per staging pass
    ("long", dict(contig_len=12000, n_ingroup=4, has_outgroup=1, depth=30.0, read_len=250, snp_density=0.03, seed=44), {}),
    # reads with mapQ below min_rmsQ contribute: their cells must leave the bit-sliced path (the rms test is per cell)
    ("lowq", dict(contig_len=9000, n_ingroup=5, has_outgroup=1, depth=24.0, snp_density=0.03, seed=45), dict(min_rmsQ=38)),
    # min_depth above the mean passing depth, min_baseQ high: coverage decided by the bit-sliced depth compare
    ("min_depth", dict(contig_len=9000, n_ingroup=2, has_outgroup=1, depth=30.0, snp_density=0.03, seed=46),
     dict(min_depth=10, min_baseQ=25)),
    # 64 samples: both halves of the 64-bit site masks, sparse coverage (empty cells, strips without records)
    ("n64_sparse", dict(contig_len=6000, n_ingroup=63, has_outgroup=1, depth=14.0, snp_density=0.05, het_frac=0.2, seed=47),
     dict(min_depth=2)),
    # Illumina-1.3 qualities (-i): the packed quality thresholds of k_planes are offset by 31
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
def test_bit_sliced_path_equals_classic_kernel_and_oracle(name, kw, pkw):
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
    assert fast.path() == 1 and classic.path() == 0, "bit-sliced path was not taken"
    orc = pbtest.OracleRun(p, fx.batch(), fx.ref(), an, wb, we)
    got, ref, want = pbtest.result_arrays(fast.res), pbtest.result_arrays(classic.res), pbtest.result_arrays(orc.res)
    assert int(want["segsites"].sum()) > 0
    assert_same(got, ref, AN_NAMES)
    assert_same(got, want, AN_NAMES)
    orc.close(); fast.close(); classic.close(); fx.close()


@pytest.mark.parametrize("name,pkw,windows", [
    # no read passes min_mapQ: every list is empty, the planes hold no passing base
    ("no_usable_reads", dict(min_mapQ=255), [(0, 3000), (3000, 6000)]),
    # windows with gaps between them, starting at an odd offset: strips straddle window edges and gap positions
    ("gaps_unaligned", {}, [(1237, 2001), (2500, 4999), (6001, 8000)]),
    # a region shorter than one strip
    ("tiny", {}, [(4111, 4130)]),
    # one-position windows
    ("single_positions", {}, [(100, 101), (5000, 5001), (8999, 9000)]),
])
def test_bit_sliced_path_region_shapes(name, pkw, windows):
    fx = pbtest.Fixture(contig_len=9000, n_ingroup=5, has_outgroup=1, depth=22.0, snp_density=0.05, het_frac=0.3, seed=61)
    p = fx.params(flags=pbtest.FLAG["OUTGROUP"], outidx=fx.n_samples - 1, **pkw)
    wb = np.array([w[0] for w in windows], dtype=np.int32); we = np.array([w[1] for w in windows], dtype=np.int32)
    an = 0
    for a in AN_NAMES:
        an |= pbtest.AN[a]
    fast = _ctx(fx, p, an, wb, we, classic=False)
    classic = _ctx(fx, p, an, wb, we, classic=True)
    assert fast.path() == 1 and classic.path() == 0, "bit-sliced path was not taken"
    orc = pbtest.OracleRun(p, fx.batch(), fx.ref(), an, wb, we)
    got, ref, want = pbtest.result_arrays(fast.res), pbtest.result_arrays(classic.res), pbtest.result_arrays(orc.res)
    assert_same(got, ref, AN_NAMES)
    assert_same(got, want, AN_NAMES)
    orc.close(); fast.close(); classic.close(); fx.close()


def test_depth_cap_binds_after_the_planes_were_built():
    """The plane pass runs before the host knows the depth bound; when the cap turns out to bind, the region falls back to
    k_pileup_call (base codes built then, quality levels taken from the plane pass) and must still equal the oracle."""
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
