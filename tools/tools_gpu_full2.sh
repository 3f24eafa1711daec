# all GPU tests on the main library, then bench of main + variants
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/tests_gpu.log 2>&1; tail -4 gpurun_out/tests_gpu.log
run() { python bench.py --steps 3 --warmup 3 --no-pruned-run --no-cpu-baseline 2>gpurun_out/bench_quick.err | tail -1 > gpurun_out/bench_$1.json; python -c "
import json,sys; d=json.load(open('gpurun_out/bench_$1.json')); print('$1', 'value %.1f' % d['value'],'e2e %.1f' % d['e2e']['value'], {k: round(v,1) for k,v in d['e2e'].get('phase_ms_per_step').items()}, 'frac %.3f' % d['roofline']['frac'], 'kernel_ms %.1f' % d['roofline']['kernel_ms'], 'terms %.3g' % d['roofline']['snp_terms_per_launch'], 'redo', d['roofline'].get('coarse_passes_repeated_exactly'))"; }
run main
for v in $(ls build/var/lib_*.so 2>/dev/null); do FSCL_B200_LIB=$PWD/$v run $(basename $v .so); done
