#!/bin/bash
# Round-2 ncu captures (run under gpurun, one GPU): launch list of the bench step and --set full captures of the four
# kernels of the resident pipeline, plus k_general at 16 reads.  Each capture runs only after the same command has exited 0
# without ncu.  Reports land in gpurun_out/ (scratch); summaries are made here with tools/ncu_summary.py / ncu_opmix.py.
set -u
A="--steps 2 --warmup 3 --no-cpu-baseline --no-secondary --no-batch96"
K='regex:k_small|k_sample|k_draw|k_window|k_segment|k_finish|k_sanitize'
timeout 300 python bench.py $A > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -3 gpurun_out/plain.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 60 --csv --log-file gpurun_out/r02_launches_bench_n1.csv python bench.py $A > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$?"
# the reports are ~45 MB each and gpurun_out/ carries 64 MiB back: summarise on the box, keep only the text
summarise() {   # <report> <units per launch> <out>
  python tools/ncu_summary.py $1 > $3 2>&1; python tools/ncu_opmix.py $1 $2 >> $3 2>&1; rm -f $1
}
for k in k_small k_window_stats_chunks k_segment_from_chunks k_sample_draws; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -o gpurun_out/r02_$k -f python bench.py $A > gpurun_out/ncu_$k.log 2>&1
  echo "$k rc=$?"
  case $k in k_small) U=229904;; k_window_stats_chunks) U=28738;; k_segment_from_chunks) U=220;; *) U=919616;; esac    # tiles of 4 draws / windows / blocks / draws
  summarise gpurun_out/r02_$k.ncu-rep $U gpurun_out/r02_ncu_$k.txt
done
G="--steps 1 --warmup 3 --no-cpu-baseline --max-reads 16 --windows 592 --subsamples 2 --depth 0.2"
timeout 300 python bench.py $G > gpurun_out/plain_g.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_general -s 3 -c 1 -o gpurun_out/r02_k_general16 -f python bench.py $G > gpurun_out/ncu_g.log 2>&1
echo "k_general rc=$?"
summarise gpurun_out/r02_k_general16.ncu-rep 1184 gpurun_out/r02_ncu_k_general16.txt      # per draw
ls -la gpurun_out/r02_ncu_*.txt
