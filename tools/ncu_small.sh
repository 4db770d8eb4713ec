#!/bin/bash
# ncu full capture of the n<=4 kernel on the bench workload (run under gpurun)
set -u
ARGS="--steps 2 --warmup 3 --no-cpu-baseline"
timeout 300 python bench.py $ARGS > gpurun_out/plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_small -s 3 -c 1 -o gpurun_out/${1:-prof_ksmall} -f python bench.py $ARGS > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?"
