#!/bin/bash
# second half of the profile suite (separate gpurun call: the .ncu-rep files are large)
set -u
mkdir -p gpurun_out
./tools/ncu_general.sh 12 2000 4 prof_kgen12
./tools/ncu_general.sh 16 600 2 prof_kgen16
./tools/bench_general.sh > gpurun_out/general_sweep.txt 2>&1
for n in 7 9 11 13 15 17 18; do ./tools/tune_general.sh $n $((n<12?3000:(n<16?1000:200))) 2 "0:0" 2>/dev/null | grep '^n=' >> gpurun_out/general_sweep.txt; done
