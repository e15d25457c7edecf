#!/usr/bin/env python
"""Runs one BASELINE.json-shaped configuration through the C ABI on cuda:0, prints stage timings and (optionally) checks
a few windows against the CPU oracle.  Used for the parity-at-scale checks and the LD profile (profiles/README.md).

  python tools/run_config.py c3      ld -o 0/-o 1, 50 kb windows, 64 samples, dense SNPs   (BASELINE configs[2])
  python tools/run_config.py c4      diverge + haplo, 32 samples                           (BASELINE configs[3], one contig)
"""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import pbtest  # noqa: E402
import popbam_b200  # noqa: E402

CONFIGS = {
    "c3": dict(fx=dict(contig_len=1000001, n_ingroup=63, has_outgroup=1, depth=20.0, snp_density=0.06, seed=303, n_threads=16),
               win=50000, an=["LD_ZNS", "LD_OMEGA", "LD_WALL"], check=["LD_ZNS", "LD_WALL"], check_windows=2),
    "c4": dict(fx=dict(contig_len=2000001, n_ingroup=31, has_outgroup=1, depth=20.0, snp_density=0.01, seed=404, n_threads=16),
               win=10000, an=["DIVERGE_IND", "DIVERGE_POP", "HAPLO_K", "HAPLO_EHHS", "HAPLO_DXY", "NUCDIV"],
               check=["DIVERGE_IND", "DIVERGE_POP", "HAPLO_K", "HAPLO_EHHS", "HAPLO_DXY", "NUCDIV"], check_windows=3),
}


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c3"
    cfg = CONFIGS[name]
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    t0 = time.time()
    fx = pbtest.Fixture(**cfg["fx"])
    print("fixture %s: %d reads, %.2f Gbases, %d samples (%.1f s)" % (name, fx.batch().n_reads, fx.aligned_bases() / 1e9, fx.n_samples, time.time() - t0))
    an = 0
    for a in cfg["an"]:
        an |= pbtest.AN[a]
    p = fx.params(flags=pbtest.FLAG["OUTGROUP"], outidx=fx.n_samples - 1)
    wb, we = pbtest.window_grid(0, fx.contig_len, cfg["win"])
    ctx = popbam_b200.Context(p)
    ctx.set_contig(0, fx.ref())
    ctx.region_begin(an, wb, we)
    ctx.push_batch(fx.batch())
    res = ctx.region_end()
    for _ in range(reps):
        ctx.relaunch()
        res = ctx.wait()
    ms = ctx.stage_times()
    got = pbtest.result_arrays(res)
    S = got["segsites"]
    print("windows %d, segsites per window: min %d median %d max %d" % (len(wb), S.min(), int(np.median(S)), S.max()))
    print("stage ms: prep %.3f  pileup %.3f  compaction %.3f  statistics %.3f" % tuple(ms))
    print("pileup: %.1f aligned Gbases/s" % (res.aligned_bases / ms[1] / 1e6))
    if an & (pbtest.AN["LD_ZNS"] | pbtest.AN["LD_OMEGA"]):
        pairs = sum(int(s) * (int(s) - 1) // 2 for s in S) * fx.n_pops
        print("LD: <= %.3e SNP pairs x pops (x2 for omega) in the statistics stage" % pairs)
    # oracle check on the first windows (the oracle's O(S^3) omega is skipped at this size)
    k = cfg["check_windows"]
    can = 0
    for a in cfg["check"]:
        can |= pbtest.AN[a]
    t0 = time.time()
    orc = pbtest.OracleRun(p, fx.batch(), fx.ref(), can, wb[:k], we[:k])
    want = pbtest.result_arrays(orc.res)
    print("oracle on %d windows: %.1f s" % (k, time.time() - t0))
    P, n = fx.n_pops, fx.n_samples
    ok = True
    for key in ("num_sites", "segsites"):
        ok &= np.array_equal(got[key][:k], want[key][:k])
    ns = int(want["seg_off"][k])
    ok &= np.array_equal(got["seg_type"][:ns], want["seg_type"][:ns]) and np.array_equal(got["seg_pos"][:ns], want["seg_pos"][:ns])
    for key, width in (("zns", P), ("wallb", P), ("wallq", P), ("piw", P), ("pib", P * P), ("hdiv", P), ("ehhs", P)):
        if len(want[key]) and len(got[key]):
            ok &= bool(np.allclose(got[key][:k * width], want[key][:k * width], rtol=1e-9, atol=0, equal_nan=True))
    for key, width in (("ld_num_snps", P), ("wall_num_snps", P), ("ind_div", n), ("pop_div", P), ("div_num_snps", P), ("nhaps", P), ("min_dxy", P * P)):
        if len(want[key]) and len(got[key]):
            ok &= bool(np.array_equal(got[key][:k * width], want[key][:k * width]))
    print("parity on the first %d windows: %s" % (k, "OK" if ok else "MISMATCH"))
    ctx.close()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
