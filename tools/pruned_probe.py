"""c2 as quoted: 1000 permutations with 20-exceedance pruning; where the time goes (run twice: memory pool warm)."""
import sys, time
sys.path.insert(0, '.')
from fscl_b200 import api, synth
from oracle import pyoracle as po
tab, opt = synth.config2()
flat = po.RefProblem(tab).flatten()
sc = api.Scanner(flat)
obs = sc.scan_chromosome(100000)
for rep in range(3):
    sc.reset_stats()
    t0 = time.perf_counter()
    p, st, _ = sc.scan_permute(obs, 1000, 100000, rng=api.RNG_PHILOX, seed=3, prune=1)
    dt = time.perf_counter() - t0
    print('rep', rep, 'seconds %.3f' % dt, {k: (round(v, 1) if isinstance(v, float) else v) for k, v in st.items()})
