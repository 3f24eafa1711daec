set -x
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -30
python - <<'PY'
import sys, time, numpy as np
sys.path.insert(0, '.')
from fscl_b200 import api, synth
from oracle import pyoracle as po
print('fp64 peak TF', api.fp64_peak_tflops())
tab, opt = synth.config2()
t0=time.time(); flat = po.RefProblem(tab).flatten(); print('prep', time.time()-t0)
sc = api.Scanner(flat)
for rep in range(3):
    sc.reset_stats(); t0=time.time(); terms, st = sc.bench_scan(100000, 0, count_terms=(rep==0)); print('fwd', time.time()-t0, terms, st['score_kernel_ms'])
for ns in (1, 4, 16):
    sc.reset_stats(); t0=time.time(); terms, st = sc.bench_scan(100000, ns, count_terms=False); dt=time.time()-t0
    print('perm slots', ns, 'wall', dt, 'kernel ms', st['score_kernel_ms'], 'permute ms', st['permute_kernel_ms'], 'scans/s', ns/(st['score_kernel_ms']*1e-3))
sc.reset_stats(); terms, st = sc.bench_scan(100000, 4, count_terms=True); print('terms 4 slots', terms, st['score_kernel_ms'])
PY
