set -x
python bench.py > gpurun_out/bench1.json 2> gpurun_out/bench1.err; echo rc=$?; tail -3 gpurun_out/bench1.err; cat gpurun_out/bench1.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null; cat gpurun_out/bench_ref.json
CMD="python bench.py --steps 1 --warmup 3 --round-perms 4 --no-cpu-baseline --no-pruned-run"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:scan_cells -s 6 -c 1 -o gpurun_out/prof_scan_r1 -f $CMD > gpurun_out/ncu2.log 2>&1
tail -5 gpurun_out/ncu2.log
