#!/usr/bin/env python
"""E tier at several sample sizes: wall time of the `popbam` command line on BAM/BAI/FASTA files (bench.py's
cli_from_bam leg with a chosen sample length), to separate process start-up from the feeder-bound rate.
Usage: python tools/cli_tier.py 1000 5000 [kb ...]"""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import bench  # noqa: E402

for kb in [int(a) for a in sys.argv[1:]] or [1000, 5000]:
    print(json.dumps({"sample_kb": kb, **bench.cli_from_bam(kb * 1000)}), flush=True)
