#!/bin/bash
# Runs on the GPU box (under gpurun): bench line, reference arm, ncu launch list and full captures.
set -u
mkdir -p gpurun_out
K='regex:k_small|k_mid|k_lanes|k_general|k_sanitize|k_sample|k_llr|k_window|k_chrom|k_bin|k_pack|k_build|k_row|k_norm|k_read|k_gather'
./tools/int_peaks > gpurun_out/INT_PEAKS.json 2> gpurun_out/int_peaks.err
mkdir -p profiles && cp gpurun_out/INT_PEAKS.json profiles/INT_PEAKS.json
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
ARGS="--steps 2 --warmup 3 --no-cpu-baseline"
timeout 300 python bench.py $ARGS > gpurun_out/plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 60 --csv --log-file gpurun_out/launches.csv python bench.py $ARGS > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
tail -c 400 gpurun_out/bench.json
# the full capture of the dominant kernel is a separate call: tools/ncu_small.sh
