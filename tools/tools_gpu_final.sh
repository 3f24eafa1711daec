# round-end evidence on one GPU: tests, bench (both arms), launch list, ncu full capture, c4 shape
set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/tests_gpu.log 2>&1; tail -3 gpurun_out/tests_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; tail -c 600 gpurun_out/bench_1gpu.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 300 gpurun_out/bench_ref.json
CMD="python bench.py --steps 1 --warmup 3 --round-perms 2 --no-cpu-baseline --no-pruned-run"
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scan_cells -s 6 -c 1 -o gpurun_out/prof_scan_final -f $CMD > gpurun_out/ncu2.log 2>&1
timeout 600 python tools/bench_genome.py --perms 200 --slots 4 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; cat gpurun_out/bench_c4.json
