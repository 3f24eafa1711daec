# all GPU tests, then a short bench of the main library and of every build/var/lib_*.so
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/tests_gpu.log 2>&1; tail -8 gpurun_out/tests_gpu.log
run() { python bench.py --steps 3 --warmup 3 --no-pruned-run --no-cpu-baseline 2>gpurun_out/bench_quick.err | tail -1 > gpurun_out/bench_$1.json; python -c "
import json,sys; d=json.load(open('gpurun_out/bench_$1.json')); print('$1', 'value %.1f' % d['value'],'e2e %.1f' % d['e2e']['value'], 'frac %.3f' % d['roofline']['frac'], 'kernel_ms %.1f' % d['roofline']['kernel_ms'], 'terms %.3g' % d['roofline']['snp_terms_per_launch'], 'redo', d['roofline'].get('coarse_passes_repeated_exactly'), 'of', d['roofline'].get('grid_evals_per_launch'), 'x steps')"; }
run main
for v in build/var/lib_*.so; do FSCL_B200_LIB=$PWD/$v run $(basename $v .so); done
