python -m pytest tests -m gpu -q -x "$@" > gpurun_out/tests_gpu.log 2>&1; tail -15 gpurun_out/tests_gpu.log
