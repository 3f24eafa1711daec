# ncu --set full of one scan_cells launch for the main lib (tag $1) and optionally a variant lib ($2 = path, tag $3)
CMD="python bench.py --steps 1 --warmup 3 --round-perms 2 --no-cpu-baseline --no-pruned-run"
ncu --set full --clock-control none --import-source on -k regex:scan_cells -s 6 -c 1 -o gpurun_out/prof_scan_$1 -f $CMD > gpurun_out/ncu_$1.log 2>&1
if [ -n "$2" ]; then
FSCL_B200_LIB=$PWD/$2 ncu --set full --clock-control none --import-source on -k regex:scan_cells -s 6 -c 1 -o gpurun_out/prof_scan_$3 -f $CMD > gpurun_out/ncu_$3.log 2>&1
fi
tail -1 gpurun_out/ncu_$1.log | cut -c1-200
