"""A/B of the scan kernel: resident permuted scans/s on c2 and c4 (FSCL_B200_LIB / FSCL_B200_PERSISTENT select the variant)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fscl_b200 import api, hostref, synth
tag = sys.argv[1] if len(sys.argv) > 1 else ""
for cfg, fn, slots in ((2, synth.config2, 16), (4, synth.config4, 4)):
    tab, opt = fn()
    flat = hostref.load_problem(table=tab, device_tables=True)
    sc = api.Scanner(flat)
    G = opt["G"]
    sc.bench_scan(G, slots)
    best = 0.0
    for rep in range(3):
        sc.reset_stats()
        t0 = time.perf_counter(); _, st = sc.bench_scan(G, slots); dt = time.perf_counter() - t0
        best = max(best, slots / (st["score_kernel_ms"] * 1e-3))
    terms, _ = sc.bench_scan(G, 1, count_terms=True)
    print(f"{tag} c{cfg}: {best:8.2f} scans/s (kernel time), {terms/1e9:.3f} G terms/scan, bound failures {sc.stats()['bound_failures']} fine {sc.stats()['fine_bound_failures']}", flush=True)
    sc.close()
