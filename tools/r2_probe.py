"""Round-2 probe: product loader + device tables + pipelined pruned test on a chosen config."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fscl_b200 import api, hostref, synth

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, default=4)
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--perms", type=int, default=1000)
ap.add_argument("--slots", type=int, default=4)
ap.add_argument("--host-tables", action="store_true")
ap.add_argument("--reps", type=int, default=2)
a = ap.parse_args()
t0 = time.time()
tab, opt = {2: synth.config2, 3: synth.config3, 4: synth.config4, 5: synth.config5}[a.config](scale=a.scale)
t_syn = time.time() - t0
t0 = time.time()
kw = {"asc_depth": opt["asc_depth"], "asc_min_freq": opt["asc_min_freq"]} if a.config == 3 else {}
flat = hostref.load_problem(table=tab, device_tables=not a.host_tables, **kw)
t_prep = time.time() - t0
G = opt["G"]
sc = api.Scanner(flat)
t0 = time.time(); obs = sc.scan_chromosome(G); t_fwd = time.time() - t0
t0 = time.time(); sc.bench_scan(G, a.slots); t_b = time.time() - t0
t0 = time.time(); sc.bench_scan(G, a.slots); t_b = time.time() - t0
out = {"workload": f"c{a.config} x {a.scale}: {len(flat['pos'])} SNPs, {len(flat['chr_start'])} chr, {len(flat['null_tab'])} classes, {len(obs)} cells, -G {G}",
       "synth_s": t_syn, "host_prep_s": t_prep, "forward_scan_s": t_fwd, "permuted_scans_per_s_resident": a.slots / t_b}
for rep in range(a.reps):
    sc.reset_stats()
    t0 = time.time()
    pts, st, _ = sc.scan_permute(obs, a.perms, G, rng=api.RNG_PHILOX, seed=4, prune=1, keep_clr=1)
    t_p = time.time() - t0
    out[f"pruned_test_{rep}"] = {"n_permute": a.perms, "seconds": t_p, "scan_equivalents": st["n_tasks"] / len(obs),
                                 "cells_finished": int(pts["permute_finished"].sum()), "rounds": st["n_rounds"],
                                 "scan_kernel_ms": st["score_kernel_ms"], "n_scored": st["n_scored"], "launches": st["kernel_launches"]}
t0 = time.time(); off, packed = sc.fetch_clr(len(obs)); t_f = time.time() - t0
t0 = time.time(); lam, p = sc.project(len(obs)); t_pr = time.time() - t0
out["fetch_clr_s"] = t_f; out["stored_clr_floats"] = int(off[-1]); out["project_s"] = t_pr
out["min_proj_p"] = float(p.min()); out["max_clr"] = float(obs["clr"].max())
print(json.dumps(out))
