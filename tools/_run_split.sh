mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/gpu_tests.log
tail -4 gpurun_out/gpu_tests.log
for n in 16 18; do
timeout 900 ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --clock-control none --import-source on -k regex:k_general -s 3 -c 1 -o gpurun_out/g$n -f python tools/sweep_sizes.py homogeneous $n > gpurun_out/ncu$n.log 2>&1
ncu -i gpurun_out/g$n.ncu-rep --page source --csv > gpurun_out/g${n}_source.csv
python tools/ncu_regions.py gpurun_out/g${n}_source.csv 100 > gpurun_out/r02_regions_homog$n.txt
gzip -f gpurun_out/g${n}_source.csv; rm -f gpurun_out/g$n.ncu-rep
done
