"""c4 as quoted (-p 10000 with pruning, one GPU) with FSCL_B200_ROUND_LOG=1: per round, the scan kernel's time and the idle
time of the scan stream before it (stderr); summary on stdout."""
import os, sys, time
os.environ["FSCL_B200_ROUND_LOG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fscl_b200 import api, hostref, synth
perms = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
tab, opt = synth.config4()
flat = hostref.load_problem(table=tab, device_tables=True)
sc = api.Scanner(flat)
G = opt["G"]
obs = sc.scan_chromosome(G)
sc.scan_permute(obs, 40, G, rng=api.RNG_PHILOX, seed=9, prune=1)       # memory pool warm
sys.stderr.write("==== timed run\n")
sc.reset_stats()
t0 = time.perf_counter()
p, st, _ = sc.scan_permute(obs, perms, G, rng=api.RNG_PHILOX, seed=4, prune=1)
dt = time.perf_counter() - t0
print("seconds %.3f" % dt, {k: (round(v, 1) if isinstance(v, float) else v) for k, v in st.items()})
