#!/bin/bash
# quick exploration of the general kernel across draw sizes (run under gpurun)
for cfg in "5 8000 16" "6 8000 16" "8 4000 8" "10 4000 4" "12 2000 4" "14 1000 2" "16 600 2"; do set -- $cfg
  timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --max-reads $1 --windows $2 --subsamples $3 --depth 0.2 > gpurun_out/bg_$1.json 2> gpurun_out/bg_$1.err
  python -c "
import json,sys; d=json.load(open('gpurun_out/bg_$1.json')); print($1, d['draws_per_step_per_gpu'], 'ms', round(d['ms_per_step'],4), 'draws/s', round(d['value']), 'int frac', round(d['int_roofline']['frac'],4))" || tail -3 gpurun_out/bg_$1.err
done
