#!/usr/bin/env python
"""Regenerates tests/golden/ from the UNMODIFIED reference (oracle/_ref, built by oracle/Makefile from
/root/reference).  Run in the build container only; the outputs are committed.

  *.txt            stdout of `popbam <argv> -f ref.fa in.bam chr1` on the seeded fixture of tests/cases.py
  kat_cells.bin    refdump cells  (errmod_cal + gl2cns + rms packing; record layout in README.md)
  kat_sites.bin    refdump sites  (clean_heterozygotes + segbase + qfilter)
  kat_tables.json  sha256 of the errmod_init tables + probe entries
"""
import hashlib
import json
import subprocess
import sys
import tempfile
from pathlib import Path

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
import numpy as np  # noqa: E402
import pbtest  # noqa: E402
from cases import CASES, FIXTURES  # noqa: E402


def main():
    pbtest.build_test_libs()
    assert pbtest.have_ref(), "reference binary missing (needs /root/reference)"
    with tempfile.TemporaryDirectory() as td:
        files = {}
        for name, kw in FIXTURES.items():
            fx = pbtest.Fixture(**kw)
            files[name] = fx.write_files(Path(td) / name)
            fx.close()
        for name, fxname, argv, _an, _pk, _ok in CASES:
            bam, fa = files[fxname]
            out = pbtest.run_ref(list(argv) + ["-f", fa, bam, "chr1"])
            (HERE / (name + ".txt")).write_text(out)
            print("%-16s %6d bytes" % (name, len(out)))
        tb = Path(td) / "tables.bin"
        subprocess.check_call([str(pbtest.REFDUMP), "tables", str(tb)])
        raw = tb.read_bytes()
        t = np.frombuffer(raw, dtype=np.float64)
        fk, beta, lhet = t[:256], t[256:256 + 64 * 65536], t[256 + 64 * 65536:]
        rng = np.random.default_rng(5)
        probes = {"beta": {}, "lhet": {}}
        for _ in range(400):
            q, n = int(rng.integers(1, 64)), int(rng.integers(1, 256))
            k = int(rng.integers(0, n + 1))
            probes["beta"]["%d" % (q << 16 | n << 8 | k)] = float(beta[q << 16 | n << 8 | k]).hex()
            probes["lhet"]["%d" % (n << 8 | k)] = float(lhet[n << 8 | k]).hex()
        json.dump({"sha256": hashlib.sha256(raw).hexdigest(), "fk": [float(x).hex() for x in fk], **probes},
                  open(HERE / "kat_tables.json", "w"), indent=0)
        subprocess.check_call([str(pbtest.REFDUMP), "cells", "101", "600", str(HERE / "kat_cells.bin")])
        subprocess.check_call([str(pbtest.REFDUMP), "sites", "202", "300", str(HERE / "kat_sites.bin")])


if __name__ == "__main__":
    main()
