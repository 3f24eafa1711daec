for pad in 0 120000; do
  FSCL_B200_SMEM_PAD=$pad python bench.py --steps 3 --warmup 3 --no-pruned-run --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('pad $pad', 'value %.1f' % d['value'],'roofline %.3f' % d['roofline']['frac'], 'kernel_ms %.1f' % d['roofline']['kernel_ms'])"
done
