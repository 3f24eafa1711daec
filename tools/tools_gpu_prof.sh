CMD="python bench.py --steps 1 --warmup 3 --round-perms 2 --no-cpu-baseline --no-pruned-run"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:scan_cells -s 6 -c 1 -o gpurun_out/prof_scan -f $CMD > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log | cut -c1-300
