#!/bin/bash
# Round profile job (run under gpurun): GPU tests, plain bench, ncu launch list, ncu --set full of the
# level-0 gate + hop launches.  Usage: tools/gpu_profile.sh <tag>
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_$TAG.log
python bench.py > $OUT/bench_$TAG.log 2>$OUT/bench_$TAG.err; echo "bench rc=$?"; cut -c1-600 $OUT/bench_$TAG.log
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > $OUT/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_launch_$TAG.log 2>&1
echo "launch list rc=$?"
$CMD > $OUT/plain2_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k "regex:hop_tc_kernel|edge_gate_tc_kernel|row_mlp_tc_kernel" -c 5 -f -o $OUT/prof_$TAG $CMD > $OUT/ncu_full_$TAG.log 2>&1
echo "full rc=$?"
