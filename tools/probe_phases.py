"""Phase latencies per grid-eval seen by thread 0 of each CTA.

Needs a library built with -DFSCL_TIMING (tools/build_variant.sh timing):
  FSCL_B200_LIB=$PWD/build/var/lib_timing.so python tools/probe_phases.py [2|4]
"""
import sys, ctypes as C, numpy as np
sys.path.insert(0, '.')
from fscl_b200 import api, hostref, synth
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 4
tab, opt = {2: synth.config2, 4: synth.config4}[cfg]()
flat = hostref.load_problem(table=tab, device_tables=True)
sc = api.Scanner(flat)
G, slots = opt["G"], (16 if cfg == 2 else 2)
n_cells = len(sc.scan_chromosome(G))
sc.bench_scan(G, slots)
buf = (C.c_ulonglong * 16)()
sc.L.fscl_debug_counters(sc._ctx, buf)
base = np.array(list(buf), dtype=np.float64)
sc.bench_scan(G, slots)
sc.L.fscl_debug_counters(sc._ctx, buf)
d = np.array(list(buf), dtype=np.float64) - base
names = ["cell bounds (per cell)", "init_scan_result", "C: direct+classify(+fine bounds)", "C: tile list", "C: walk (thread 0)", "C: wait+reduce sync", "C: combine",
         "F: direct+classify(+fine bounds)", "F: tile list", "F: walk (thread 0)", "F: wait+reduce sync", "F: combine", "argmax coarse", "bookkeeping"]
n_cta = float(slots * n_cells)
tot = d[2:16].sum()
print("c%d: cycles per CTA: %.0f  (per eval %.0f)" % (cfg, tot / n_cta, tot / n_cta / 12))
for i, nm in enumerate(names):
    print(f"{nm:34s} {d[2+i]/n_cta:10.0f} cycles/CTA  {100*d[2+i]/tot:5.1f}%")
