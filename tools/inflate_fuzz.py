#!/usr/bin/env python
"""Corrupted DEFLATE streams through the feeder's decoder under AddressSanitizer + UBSan (CPU only): every stream must be
rejected or decoded without a sanitizer report -- the decoder's bounds tests are what stand between a damaged BAM and the
read batches.  Usage: python tools/inflate_fuzz.py [n]   (builds /tmp/popbam_asan from popbam_b200/csrc)"""
import os
import re
import subprocess
import sys
import zlib
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests"))
sys.path.insert(0, str(ROOT))
from test_host_feeder import _bgzf_block  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 600
src = ROOT / "popbam_b200" / "csrc"
bld = ROOT / "popbam_b200" / "_build"
exe = "/tmp/popbam_asan"
subprocess.run(["g++", "-std=c++17", "-g", "-O1", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-o", exe,
                str(src / "popbam_main.cpp"), str(src / "pb_bamio.cpp"), str(src / "pb_inflate.cpp"), "-L" + str(bld), "-lpopbam_b200",
                "-Wl,-rpath," + str(bld), "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64", "-lpthread"], check=True)
rng = np.random.default_rng(123)
pay = [rng.choice(np.frombuffer(b"!5?DI", dtype=np.uint8), 3000).tobytes(), b"ACGTTGCA" * 500, rng.integers(0, 256, 2000, dtype=np.uint8).tobytes(),
       bytes(rng.integers(0, 4, 5000, dtype=np.uint8)), bytes(40000)]
os.makedirs("/tmp/fz", exist_ok=True)
reports = rejected = decoded = 0
for it in range(n):
    pl = pay[it % len(pay)]
    blk = bytearray(_bgzf_block(pl, [1, 6, 9][it % 3], [zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY][(it // 5) % 3]))
    for _ in range(int(rng.integers(1, 4))):                     # damage 1-3 bytes of the DEFLATE payload
        k = int(rng.integers(18, len(blk) - 8))
        blk[k] ^= int(rng.integers(1, 256))
    if it % 7 == 0:                                             # or cut the payload short (the block size field still fits)
        cut = int(rng.integers(1, max(2, (len(blk) - 26) // 2)))
        blk = blk[:len(blk) - 8 - cut] + blk[-8:]
        blk[16:18] = (len(blk) - 1).to_bytes(2, "little")
    open("/tmp/fz/x.bgzf", "wb").write(bytes(blk) + _bgzf_block(b"", 6, zlib.Z_DEFAULT_STRATEGY))
    r = subprocess.run([exe, "_inflate", "/tmp/fz/x.bgzf", "/tmp/fz/out"], stderr=subprocess.PIPE, text=True)
    if "Sanitizer" in r.stderr or re.search(r":\d+:\d+: runtime error", r.stderr):
        reports += 1
        print(r.stderr[:800])
        break
    if r.returncode != 0:
        rejected += 1
    else:
        decoded += 1
print("%d damaged streams: %d sanitizer reports, %d rejected, %d decoded to something without an error" % (n, reports, rejected, decoded))
sys.exit(1 if reports else 0)
