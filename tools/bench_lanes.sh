#!/bin/bash
# draw sizes served by k_lanes, with and without it (run under gpurun): tools/bench_lanes.sh [model]
MODEL=${1:-homogeneous}
for cfg in "5 8000 16" "6 8000 16" "7 6000 8" "8 4000 8" "9 4000 4" "10 4000 4"; do set -- $cfg
  for off in 0 1; do
    if [ $off = 1 ]; then export LDPGTA_NO_LANES=1; else unset LDPGTA_NO_LANES; fi
    timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --model $MODEL --max-reads $1 --windows $2 --subsamples $3 --depth 0.2 > gpurun_out/bl_$1_$off.json 2> gpurun_out/bl_$1_$off.err
    python -c "
import json,sys; d=json.load(open('gpurun_out/bl_$1_$off.json')); print('$MODEL n=$1', 'table-kernel' if $off else 'k_lanes     ', d['draws_per_step_per_gpu'], 'ms', round(d['ms_per_step'],4), 'draws/s %.3e' % d['value'], 'int frac', round(d['int_roofline']['frac'],4))" || tail -3 gpurun_out/bl_$1_$off.err
  done
done
