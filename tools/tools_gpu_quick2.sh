python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "small50 or dense or ascnan" > gpurun_out/tq.log 2>&1; tail -3 gpurun_out/tq.log
python bench.py --steps 3 --warmup 3 --no-pruned-run --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('value',d['value'],'e2e',d['e2e']['value'],'roofline',d['roofline']['frac'],d['roofline']['kernel_ms'], d['clocks'])"
