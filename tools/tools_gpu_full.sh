# all GPU tests, then a short bench of the main library
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/tests_gpu.log 2>&1; tail -15 gpurun_out/tests_gpu.log
python bench.py --steps 3 --warmup 3 --no-pruned-run --no-cpu-baseline 2>gpurun_out/bench_quick.err | tail -1 > gpurun_out/bench_quick.json; cat gpurun_out/bench_quick.json
