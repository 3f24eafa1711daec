"""Phase latencies per grid-eval seen by thread 0 of each CTA.

Needs a library built with -DFSCL_TIMING:
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -shared -DFSCL_TIMING \\
       -o build/var/lib_timing.so fscl_b200/csrc/fscl_{kernels,host,tables}.cu
  FSCL_B200_LIB=$PWD/build/var/lib_timing.so python tools/probe_phases.py
"""
import sys, ctypes as C, numpy as np
sys.path.insert(0, '.')
from fscl_b200 import api, synth
from oracle import pyoracle as po
tab, opt = synth.config2()
flat = po.RefProblem(tab).flatten()
sc = api.Scanner(flat)
sc.bench_scan(100000, 16)
buf = (C.c_ulonglong * 16)()
sc.L.fscl_debug_counters(sc._ctx, buf)
base = np.array(list(buf), dtype=np.float64)
sc.bench_scan(100000, 16)
sc.L.fscl_debug_counters(sc._ctx, buf)
d = np.array(list(buf), dtype=np.float64) - base
names = ["cell bounds (per cell)", "init_scan_result", "C: direct+classify", "C: tile list", "C: walk (thread 0)", "C: wait+reduce sync", "C: combine",
         "F: direct+classify", "F: tile list", "F: walk (thread 0)", "F: wait+reduce sync", "F: combine", "argmax coarse", "bookkeeping"]
n_cta = 16000.0
tot = d[2:16].sum()
print("cycles per CTA: %.0f  (per eval %.0f)" % (tot / n_cta, tot / n_cta / 12))
for i, nm in enumerate(names):
    print(f"{nm:28s} {d[2+i]/n_cta:10.0f} cycles/CTA  {100*d[2+i]/tot:5.1f}%")
