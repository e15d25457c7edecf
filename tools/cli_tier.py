#!/usr/bin/env python
"""E tier at several sample sizes: wall time of the `popbam` command line on BAM/BAI/FASTA files (bench.py's
cli_from_bam leg with a chosen sample length), to separate process start-up from the feeder-bound rate.
Usage: python tools/cli_tier.py [config] 1000 5000 [kb ...]      (config: c1..c5, default c2)"""
import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import bench  # noqa: E402

argv = sys.argv[1:]
cfg = argv.pop(0) if argv and argv[0] in bench.CONFIGS else "c2"
args = argparse.Namespace(cfg=bench.CONFIGS[cfg], config=cfg)
for kb in [int(a) for a in argv] or [1000, 5000]:
    print(json.dumps({"sample_kb": kb, **bench.cli_from_bam(args, kb * 1000)}), flush=True)
