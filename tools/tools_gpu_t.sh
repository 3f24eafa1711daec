# selected GPU tests (args: pytest -k expression), output tail
timeout 1500 python -m pytest tests -m gpu -x -q -k "$1" > gpurun_out/tests_sel.log 2>&1; tail -25 gpurun_out/tests_sel.log
