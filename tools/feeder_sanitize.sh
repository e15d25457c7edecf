#!/bin/bash
# ThreadSanitizer and AddressSanitizer + UBSan runs of the command line's host side (no GPU needed): the feeder threads
# (`popbam _feed`: shards, pieces, decode threads, batch pool, workers, printer), the single fetch and the index builder.
# usage: bash tools/feeder_sanitize.sh [outdir]      (logs: <outdir>/feeder_tsan.txt, feeder_asan.txt)
set -e
root=$(cd "$(dirname "$0")/.." && pwd)
out=${1:-$root/profiles}
src=$root/popbam_b200/csrc
bld=$root/popbam_b200/_build
tmp=$(mktemp -d)
trap 'rm -rf "$tmp"' EXIT
python - "$tmp" <<'PY'
import sys
sys.path.insert(0, "tests")
import pbtest
# 0.4 Mb, 8 samples in 2 x 2 read groups, deletions / clips / N operations: ~50 shards of 8 windows, several pieces in the no-window run
fx = pbtest.Fixture(contig_len=400000, n_ingroup=7, has_outgroup=1, rg_per_sample=2, depth=12.0, snp_density=0.02, edge_mode=1, seed=5)
print(*fx.write_files(sys.argv[1] + "/s"))
PY
bam=$tmp/s.bam; fa=$tmp/s.fa
for kind in tsan asan; do
  if [ $kind = tsan ]; then flags="-fsanitize=thread"; else flags="-fsanitize=address,undefined -fno-sanitize-recover=undefined"; fi
  g++ -std=c++17 -g -O1 $flags -o $tmp/popbam_$kind $src/popbam_main.cpp $src/pb_bamio.cpp $src/pb_inflate.cpp -L$bld -lpopbam_b200 \
      -Wl,-rpath,$bld -L/usr/local/cuda/lib64 -Wl,-rpath,/usr/local/cuda/lib64 -lpthread
  log=$out/r2_feeder_$kind.txt
  {
    echo "# $kind build of popbam_main.cpp + pb_bamio.cpp + pb_inflate.cpp ($flags), $(g++ --version | head -1)"
    echo "# BAM: $(stat -c %s $bam) bytes, 400 kb contig, 8 samples x 2 read groups, edge-mode CIGARs"
    for cmd in "index $bam" "_fetch $bam chr1:100001-300000 $tmp/f.bin" "_fetch $bam chr1:100001-300000 $tmp/g.bin 9" \
               "_feed -f $fa -w 1 --shard-mb 0.008 --threads 16 --gpus 4 $bam chr1" "_feed -f $fa --threads 12 --gpus 2 $bam chr1" \
               "_feed -f $fa -w 5 --shard-mb 0.02 --threads 3 --gpus 1 $bam chr1:50001-350000"; do
      echo "\$ popbam ${cmd//$tmp\//}"
      set +e
      TSAN_OPTIONS="halt_on_error=0" ASAN_OPTIONS="detect_leaks=1" $tmp/popbam_$kind $cmd > $tmp/stdout.txt 2> $tmp/stderr.txt
      rc=$?
      set -e
      echo "exit code $rc, $(wc -l < $tmp/stdout.txt) lines of output, sanitizer reports: $(grep -c -E 'WARNING: ThreadSanitizer|ERROR: AddressSanitizer|runtime error:|ERROR: LeakSanitizer' $tmp/stderr.txt || true)"
      grep -v "^\[popbam index\]" $tmp/stderr.txt | head -40
    done
    cmp $tmp/f.bin $tmp/g.bin && echo "one fetch == nine pieces"
  } > $log 2>&1
  tail -3 $log
done
