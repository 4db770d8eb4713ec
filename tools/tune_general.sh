#!/bin/bash
# usage: tune_general.sh n windows subsamples "NW:KB NW:KB ..." [TH]
n=$1; w=$2; s=$3
for cfg in $4; do nw=${cfg%%:*}; kb=${cfg##*:}
  if [ "$nw" = 0 ]; then unset LDPGTA_PLAN_NW LDPGTA_PLAN_KB; EV=""; else EV="LDPGTA_PLAN_NW=$nw LDPGTA_PLAN_KB=$kb"; fi
  env $EV ${5:+LDPGTA_PLAN_TH=$5} timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --max-reads $n --windows $w --subsamples $s --depth 0.2 > gpurun_out/tg.json 2> gpurun_out/tg.err
  python -c "
import json,sys; d=json.load(open('gpurun_out/tg.json')); print('n=$n nw=$nw kb=$kb th=${5:-default}', 'ms', round(d['ms_per_step'],4), 'draws/s', round(d['value']), 'int frac', round(d['int_roofline']['frac'],4))" || tail -2 gpurun_out/tg.err
done
