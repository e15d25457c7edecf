"""Parity at a size the CPU oracle cannot cover in seconds, through size-independent properties (GPU box):
  * sharding invariance -- a 1 Mb region (C1-shaped: 10 + 1 samples, 20x) processed in one piece equals the
    concatenation of three region shards fed only with their own overlapping reads (what the multi-GPU driver does);
  * idempotence -- re-running the pipeline on the resident reads reproduces every result bit for bit;
  * spot checks -- randomly chosen windows agree with the CPU oracle."""
import ctypes as C

import numpy as np
import pytest

import pbtest
import popbam_b200
from test_gpu_parity import _slice_batch, assert_same

pytestmark = pytest.mark.gpu

AN = pbtest.AN["NUCDIV"] | pbtest.AN["SFS"] | pbtest.AN["LD_ZNS"] | pbtest.AN["DIVERGE_POP"] | pbtest.AN["HAPLO_K"]
NAMES = ["NUCDIV", "SFS", "LD_ZNS", "DIVERGE_POP", "HAPLO_K"]


@pytest.fixture(scope="module")
def big():
    fx = pbtest.Fixture(contig_len=1000001, n_ingroup=10, has_outgroup=1, depth=20.0, snp_density=0.01, seed=1, n_threads=8)
    yield fx
    fx.close()


def _run(ctx, fx, wb, we, batch):
    ctx.region_begin(AN, wb, we)
    ctx.push_batch(batch)
    ctx.region_end()
    return pbtest.result_arrays(ctx.res)


def test_sharding_invariance_idempotence_and_spot_checks(big):
    fx = big
    p = fx.params(flags=pbtest.FLAG["OUTGROUP"], outidx=fx.n_samples - 1)
    wb, we = pbtest.window_grid(0, fx.contig_len, 10000)
    assert len(wb) == 100
    ctx = popbam_b200.Context(p)
    ctx.set_contig(0, fx.ref())
    b = fx.batch()
    whole = _run(ctx, fx, wb, we, b)
    assert int(whole["num_sites"].sum()) > 900000 and int(whole["seg_off"][-1]) > 5000
    # idempotence on resident reads
    ctx.relaunch(); ctx.wait()
    assert_same(pbtest.result_arrays(ctx.res), whole, NAMES)
    # three shards, each fed only the reads overlapping its span (bam_fetch semantics: start < end, reference end > begin)
    pos = np.ctypeslib.as_array(b.pos, (b.n_reads,))
    P, n = fx.n_pops, fx.n_samples
    cuts = [0, 33, 71, 100]
    keep = []
    for a, z in zip(cuts[:-1], cuts[1:]):
        lo = int(np.searchsorted(pos, wb[a] - 200, side="left"))      # 200 > any read's reference span here
        hi = int(np.searchsorted(pos, we[z - 1], side="left"))
        part = _run(ctx, fx, wb[a:z], we[a:z], _slice_batch(b, lo, hi, keep))
        assert np.array_equal(part["num_sites"], whole["num_sites"][a:z])
        assert np.array_equal(part["segsites"], whole["segsites"][a:z])
        s0, s1 = whole["seg_off"][a], whole["seg_off"][z]
        assert np.array_equal(part["seg_type"], whole["seg_type"][s0:s1])
        assert np.array_equal(part["seg_pos"], whole["seg_pos"][s0:s1])
        for key, w in (("piw", P), ("pib", P * P), ("td", P), ("fwh", P), ("zns", P), ("hdiv", P)):
            assert np.allclose(part[key], whole[key][a * w:z * w], rtol=1e-9, atol=0, equal_nan=True), key
        for key, w in (("sfs_num_snps", P), ("ld_num_snps", P), ("pop_div", P), ("div_num_snps", P), ("nhaps", P)):
            assert np.array_equal(part[key], whole[key][a * w:z * w]), key
    # oracle spot checks on three windows
    for w in (7, 52, 99):
        orc = pbtest.OracleRun(p, b, fx.ref(), AN, wb[w:w + 1], we[w:w + 1])
        want = pbtest.result_arrays(orc.res)
        assert want["num_sites"][0] == whole["num_sites"][w]
        s0, s1 = whole["seg_off"][w], whole["seg_off"][w + 1]
        assert np.array_equal(want["seg_type"], whole["seg_type"][s0:s1])
        for key in ("piw", "td", "fwh", "zns", "hdiv"):
            assert np.allclose(want[key], whole[key][w * P:(w + 1) * P], rtol=1e-9, atol=0, equal_nan=True), key
        orc.close()
    ctx.close()
