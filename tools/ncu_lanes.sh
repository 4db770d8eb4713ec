#!/bin/bash
# ncu full capture of k_lanes: tools/ncu_lanes.sh <n> <windows> <subsamples> <outname> [model]
set -u
ARGS="--steps 1 --warmup 3 --no-cpu-baseline --model ${5:-homogeneous} --max-reads $1 --windows $2 --subsamples $3 --depth 0.2"
timeout 300 python bench.py $ARGS > gpurun_out/plain_l.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_lanes -s 3 -c 1 -o gpurun_out/$4 -f python bench.py $ARGS > gpurun_out/ncu_l.log 2>&1
echo "ncu rc=$?"
