# A/B: main lib vs every build/var/lib_*.so (bench only), then an ncu --set full capture of the main lib's scan kernel
run() { python bench.py --steps 3 --warmup 3 --no-pruned-run --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', 'value %.1f' % d['value'],'roofline %.3f' % d['roofline']['frac'], 'kernel_ms %.1f' % d['roofline']['kernel_ms'])"; }
run main
for v in build/var/lib_*.so; do FSCL_B200_LIB=$PWD/$v run $v; done
if [ -n "$NCU_TAG" ]; then
CMD="python bench.py --steps 1 --warmup 3 --round-perms 2 --no-cpu-baseline --no-pruned-run"
ncu --set full --clock-control none --import-source on -k regex:scan_cells -s 6 -c 1 -o gpurun_out/prof_scan_$NCU_TAG -f $CMD > gpurun_out/ncu_$NCU_TAG.log 2>&1
tail -2 gpurun_out/ncu_$NCU_TAG.log | cut -c1-200
fi
