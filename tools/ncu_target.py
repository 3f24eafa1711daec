"""Target for ncu captures: a few resident permuted scans of c2 / c4 (every kernel of a round appears once per call)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fscl_b200 import api, hostref, synth
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 4
slots = int(sys.argv[2]) if len(sys.argv) > 2 else 1
tab, opt = {2: synth.config2, 4: synth.config4}[cfg]()
flat = hostref.load_problem(table=tab, device_tables=True)
if len(sys.argv) > 3:        # the ascertainment kernel of the table build as well
    api.build_tables(flat["depths"], flat["fsp"], asc_depth=20, asc_min_freq=1)
sc = api.Scanner(flat)
G = opt["G"]
obs = sc.scan_chromosome(G)
for rep in range(3):
    sc.bench_scan(G, slots)
if len(sys.argv) > 3:        # a short pruned test: gen_blocks / permute / bucket / taskgen / replay / compact / pack / project
    pts, st, _ = sc.scan_permute(obs, int(sys.argv[3]), G, rng=api.RNG_PHILOX, seed=3, keep_clr=1)
    sc.fetch_clr(len(obs)); sc.project(len(obs))
    print(st)
