import sys, time
sys.path.insert(0, '.')
import numpy as np
from fscl_b200 import api
from bench import build_problem, G
tab, rp, flat = build_problem()
sc = api.Scanner(flat)
obs = sc.scan_chromosome(G)
for rep in range(2):
    sc.reset_stats()
    t0 = time.perf_counter()
    p2, st, _ = sc.scan_permute(obs, 1000, G, rng=api.RNG_PHILOX, seed=3, prune=1)
    dt = time.perf_counter() - t0
    print(rep, 'wall %.3f s' % dt, 'score_ms %.1f' % st['score_kernel_ms'], 'perm_ms %.1f' % st['permute_kernel_ms'], 'rounds', st['n_rounds'], 'tasks', st['n_tasks'], 'scored', st['n_scored'], 'finished', int(p2['permute_finished'].sum()))
