#!/bin/bash
# one measurement round on the GPU box: fast-path parity tests, launch-shape sweep, ncu capture of k_pile_reads, launch list
# usage (under gpurun): bash tools/gpu_iter.sh <outdir-name> [shapes...]
out=gpurun_out/$1; shift
mkdir -p $out
[ -n "$NO_TESTS" ] || timeout 600 python -m pytest tests/test_gpu_fastpath.py -x -q -m gpu 2>&1 | tail -15
B="python bench.py --steps 1 --warmup 1 --contig-mb 2.3 --inflight 1 --no-cpu-baseline --cli-sample-kb 0"
for cfg in "$@"; do
  [ "$cfg" = auto ] && { $B 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('auto', j['ms_per_step'], j['stage_ms_per_shard'])"; continue; }
  POPBAM_B200_PILE=$cfg $B 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('$cfg', j['ms_per_step'], j['stage_ms_per_shard'])"
done
ncu --set full --clock-control none --import-source on -k regex:k_pile_reads -c 1 -s 2 -o $out/pile $B --no-verify > $out/ncu.log 2>&1
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 80 --csv --log-file $out/launches.csv $B --no-verify > /dev/null 2>&1
