# ncu --set full of one scan_cells launch (main lib), tag = $1
CMD="python bench.py --steps 1 --warmup 3 --round-perms 2 --no-cpu-baseline --no-pruned-run"
ncu --set full --clock-control none --import-source on -k regex:scan_cells -s 6 -c 1 -o gpurun_out/prof_scan_$1 -f $CMD > gpurun_out/ncu_$1.log 2>&1
tail -2 gpurun_out/ncu_$1.log | cut -c1-200
