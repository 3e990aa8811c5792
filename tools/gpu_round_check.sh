#!/bin/bash
# Round check (run under gpurun): all GPU tests, the headline bench, both training benches, and the ncu launch list
# of the headline command.  Usage: tools/gpu_round_check.sh <tag>
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_$TAG.log
python bench.py > $OUT/bench_$TAG.log 2>$OUT/bench_$TAG.err; echo "bench rc=$?"; cut -c1-300 $OUT/bench_$TAG.log
python bench.py --workload cfg2-train > $OUT/bench_cfg2train_$TAG.log 2>$OUT/bench_cfg2train_$TAG.err; echo "cfg2-train rc=$?"
python bench.py --workload cfg5-train > $OUT/bench_cfg5train_$TAG.log 2>$OUT/bench_cfg5train_$TAG.err; echo "cfg5-train rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke_$TAG.log
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-training-extra"
$CMD > $OUT/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_launch_$TAG.log 2>&1
echo "launch list rc=$?"
