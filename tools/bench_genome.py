"""Other BASELINE.json shapes on one GPU: --config 3 (configs[2]: array-style input, ascertainment K=2 of M=20, 5 depths,
30 % folded), --config 4 (BASELINE.json configs[3]: 3M SNPs, 22 autosomes, -G 100000) or --config 5
(configs[4]: 20M SNPs, depths 80..100, 5 % folded, -G 10000): forward scan, permuted scans/s with inputs resident, and a
pruned permutation test of --perms permutations."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fscl_b200 import api, synth
from oracle import pyoracle as po          # input preparation by the reference's host code only

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, default=4)
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--perms", type=int, default=200)
ap.add_argument("--slots", type=int, default=4)
a = ap.parse_args()
t0 = time.time()
tab, opt = {3: synth.config3, 4: synth.config4, 5: synth.config5}[a.config](scale=a.scale)
flat = po.RefProblem(tab, **({"asc_depth": opt["asc_depth"], "asc_min_freq": opt["asc_min_freq"]} if a.config == 3 else {})).flatten()
t_prep = time.time() - t0
G = opt["G"]
sc = api.Scanner(flat)
t0 = time.time(); obs = sc.scan_chromosome(G); t_fwd = time.time() - t0
sc.reset_stats()
t0 = time.time(); sc.bench_scan(G, a.slots); t_b = time.time() - t0
st = sc.stats()
t0 = time.time(); sc.bench_scan(G, a.slots); t_b = time.time() - t0
t0 = time.time(); pts, st2, _ = sc.scan_permute(obs, a.perms, G, rng=api.RNG_PHILOX, seed=4, prune=1); t_p = time.time() - t0
print(json.dumps({"workload": f"c{a.config} x {a.scale}: {len(flat['pos'])} SNPs, {len(flat['chr_start'])} chromosomes, {len(flat['null_tab'])} classes, {len(obs)} cells, -G {G}",
                  "host_prep_s": t_prep, "forward_scan_s": t_fwd, "permuted_scans_per_s_resident": a.slots / t_b,
                  "pruned_test": {"n_permute": a.perms, "seconds": t_p, "scan_equivalents": st2["n_tasks"] / len(obs),
                                  "cells_finished": int(pts["permute_finished"].sum()), "rounds": st2["n_rounds"],
                                  "scan_kernel_ms": st2["score_kernel_ms"], "permute_kernels_ms": st2["permute_kernel_ms"]},
                  "max_clr": float(obs["clr"].max())}))
