#!/usr/bin/env python
"""Multi-GPU check of the command line: the same BAM through `popbam --gpus 1` and `popbam --gpus N` (region shards dealt
round-robin to the devices, rows printed in window order) must give byte-identical output.  Usage: cli_multi_gpu_check.py N"""
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import pbtest  # noqa: E402
import popbam_b200  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    exe = popbam_b200.capi.PKG / "_build" / "popbam"
    fx = pbtest.Fixture(contig_len=400001, n_ingroup=10, has_outgroup=1, depth=20.0, snp_density=0.01, seed=91, n_threads=8)
    with tempfile.TemporaryDirectory() as td:
        bam, fa = fx.write_files(Path(td) / "m")
        outs = []
        for g in (1, n):
            for cmd in (["nucdiv", "-w", "10"], ["sfs", "-w", "10", "-p", "og"], ["ld", "-w", "10", "-o", "1"]):
                t0 = time.time()
                r = subprocess.run([str(exe)] + cmd + ["--gpus", str(g), "--shard-mb", "0.05", "-f", fa, bam, "chr1"],
                                   stdout=subprocess.PIPE, stderr=subprocess.PIPE)
                assert r.returncode == 0, r.stderr.decode()
                outs.append((g, cmd[0], r.stdout, time.time() - t0))
        k = len(outs) // 2
        for a, b in zip(outs[:k], outs[k:]):
            same = a[2] == b[2]
            print("%-7s rows %3d  --gpus 1: %.2f s  --gpus %d: %.2f s  identical: %s" % (a[1], a[2].count(b"\n"), a[3], n, b[3], same))
            assert same
    print("multi-GPU command line OK")


if __name__ == "__main__":
    main()
