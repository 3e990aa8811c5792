#!/bin/bash
# Round-2 profile capture (run under gpurun, ONE GPU): launch list of the headline command and ncu --set full captures of
# the level-0 launches of the three dominant kernels.  Usage: tools/gpu_profile_r02.sh <tag>
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-training-extra"
$CMD > $OUT/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/plain_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_launch_$TAG.log 2>&1
echo "launch list rc=$?"
# level-0 launches of the second step: the 11th gate, the 29th tensor-core hop, the 12th row-MLP launch (static encoder)
ncu --set full --clock-control none --import-source on -k regex:edge_gate_tc16_kernel -s 10 -c 1 -o $OUT/prof_gate_$TAG $CMD > $OUT/ncu_gate_$TAG.log 2>&1; echo "gate rc=$?"
ncu --set full --clock-control none --import-source on -k regex:hop_tc_kernel -s 28 -c 1 -o $OUT/prof_hop_$TAG $CMD > $OUT/ncu_hop_$TAG.log 2>&1; echo "hop rc=$?"
ncu --set full --clock-control none --import-source on -k regex:row_mlp_tc_kernel -s 11 -c 3 -o $OUT/prof_rowmlp_$TAG $CMD > $OUT/ncu_rowmlp_$TAG.log 2>&1; echo "rowmlp rc=$?"
ls -la $OUT/*.ncu-rep
