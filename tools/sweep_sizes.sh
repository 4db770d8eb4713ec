#!/bin/bash
# every draw size 2..18 through bench.py (one GPU): kernel, draws/s, fraction of the measured AND+POPC roofline.
# usage (under gpurun): tools/sweep_sizes.sh [model] > gpurun_out/sweep.txt
MODEL=${1:-homogeneous}
declare -A CFG=([2]="20000 32" [3]="20000 32" [4]="28738 32" [5]="8000 32" [6]="8000 32" [7]="6000 32" [8]="4000 32" [9]="4000 16" [10]="4000 8"
                [11]="2000 8" [12]="2000 4" [13]="1480 2" [14]="1036 2" [15]="592 2" [16]="592 2" [17]="296 2" [18]="296 1")      # 13+: one draw per CTA, whole waves of 148
for n in ${SWEEP_N:-2 3 4 5 6 7 8 9 10 11 12 13 14 15 16 17 18}; do set -- ${CFG[$n]}
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --model $MODEL --max-reads $n --windows $1 --subsamples $2 --depth 0.2 > gpurun_out/sw_$n.json 2> gpurun_out/sw_$n.err
  python -c "
import json; d=json.load(open('gpurun_out/sw_$n.json'))
print('$MODEL n=%2d  %-18s draws/step %7d  ms %8.4f  draws/s %.3e  popc-roofline frac %.3f  hbm frac %.3f' % ($n, d['roofline']['kernel'], d['draws_per_step_per_gpu'], d['ms_per_step'], d['value'], d['int_roofline']['frac'], d['roofline']['frac']))" || tail -3 gpurun_out/sw_$n.err
done
