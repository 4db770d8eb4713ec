#!/bin/bash
# draw sizes served by k_lanes, with and without it (run under gpurun): tools/bench_lanes.sh [model] [sizes...]
MODEL=${1:-homogeneous}; shift
SIZES=${@:-5 6 7 8 9 10}
declare -A CFG=([5]="8000 32" [6]="8000 32" [7]="6000 32" [8]="4000 32" [9]="4000 16" [10]="4000 8" [11]="2000 8" [12]="2000 4")
for n in $SIZES; do set -- ${CFG[$n]}
  for off in 0 1; do
    if [ $off = 1 ]; then export LDPGTA_NO_LANES=1; else unset LDPGTA_NO_LANES; fi
    timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --model $MODEL --max-reads $n --windows $1 --subsamples $2 --depth 0.2 > gpurun_out/bl_${n}_$off.json 2> gpurun_out/bl_${n}_$off.err
    python -c "
import json,sys; d=json.load(open('gpurun_out/bl_${n}_$off.json')); print('$MODEL n=$n', 'table-kernel' if $off else 'k_lanes     ', d['draws_per_step_per_gpu'], 'ms', round(d['ms_per_step'],4), 'draws/s %.3e' % d['value'], 'int frac', round(d['int_roofline']['frac'],4))" || tail -3 gpurun_out/bl_${n}_$off.err
  done
done
