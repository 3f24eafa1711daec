"""Where the evaluated SNP-terms are: by pass, code path and alpha candidate.

Needs a library built with -DFSCL_HIST (tools/build_variant.sh hist):
  FSCL_B200_LIB=$PWD/build/var/lib_hist.so python tools/probe_terms.py [--config 2|4]
"""
import argparse, ctypes as C, sys
import numpy as np
sys.path.insert(0, '.')
from fscl_b200 import api, hostref, synth
ap = argparse.ArgumentParser(); ap.add_argument("--config", type=int, default=4); a = ap.parse_args()
tab, opt = {2: synth.config2, 4: synth.config4}[a.config]()
flat = hostref.load_problem(table=tab, device_tables=True)
sc = api.Scanner(flat)
G = opt["G"]
n_cells = len(sc.scan_chromosome(G))
terms, _ = sc.bench_scan(G, 1, count_terms=True)
buf = (C.c_ulonglong * 128)()
sc.L.fscl_debug_histogram(sc._ctx, buf)
h = np.array(list(buf), dtype=np.float64).reshape(2, 4, 16) / (n_cells * 12)
print(f"c{a.config}: {terms / (n_cells * 12):.0f} terms per grid-eval")
for p, pn in enumerate(("coarse", "fine")):
    for c, cn in enumerate(("direct", "covering", "crossing", "behind-bound")):
        print(f"{pn:7s} {cn:13s} total {h[p, c].sum():8.0f}  by candidate:", " ".join(f"{v:6.0f}" for v in h[p, c, :15]))
