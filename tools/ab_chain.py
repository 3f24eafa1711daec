"""End-point sharing on / off at different launch sizes: resident permuted scans/s by kernel time (FSCL_B200_LIB selects the library)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fscl_b200 import api, hostref, synth
tag = sys.argv[1] if len(sys.argv) > 1 else ""
for cfg, fn, slot_list in ((2, synth.config2, (4, 16, 64)), (4, synth.config4, (1, 4))):
    tab, opt = fn()
    flat = hostref.load_problem(table=tab, device_tables=True)
    sc = api.Scanner(flat)
    G = opt["G"]
    for slots in slot_list:
        for on in (False, True):
            sc.set_endpoint_sharing(on)
            sc.bench_scan(G, slots)
            best = 0.0
            for rep in range(3):
                sc.reset_stats()
                _, st = sc.bench_scan(G, slots)
                best = max(best, slots / (st["score_kernel_ms"] * 1e-3))
            print(f"{tag} c{cfg} slots {slots:3d} sharing {'on ' if on else 'off'}: {best:8.2f} scans/s", flush=True)
    sc.close()
