"""Stall samples / instructions per section of scan_cells_kernel from `ncu --page source --print-source cuda,sass --csv`."""
import csv, sys, os
rows = list(csv.reader(open(sys.argv[1])))
agg = {}; hdr = None
for r in rows:
    if len(r) > 8 and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < 8 or r[2] != "-": continue
    try: ln = int(r[0]); s = int(r[6] or 0); e = int(r[7] or 0)
    except ValueError: continue
    a = agg.setdefault(ln, [0, 0]); a[0] += s; a[1] += e
ts = sum(a[0] for a in agg.values()); te = sum(a[1] for a in agg.values())
src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "fscl_b200/csrc/fscl_kernels.cu")).read().split("\n")
def find(pat): return next(i + 1 for i, l in enumerate(src) if pat in l)
marks = [("helpers (logt, row loads, batch FMAs)", 1), ("cell bounds", find("__device__ void dev_cell_bounds")),
         ("phase head", find("template <int NA, bool COUNT, bool WITHN>\n".strip())), ("direct candidates", find("candidates with a narrow cut are summed")),
         ("classify tiles", find("classify the tiles of the window once per pass")), ("tile list + null sums", find("tiles that need the per-site walk, in order")),
         ("walk: records, distances, log table", find("A warp owns chunks")), ("walk: covering candidates", find("alphas whose cut covers the whole tile: no predicate")),
         ("walk: crossing candidates", find("alphas whose cut crosses the tile (or a tile cut")), ("reduction", find("fixed-tree reduction: deterministic")),
         ("search_maxpos control", find("search_maxpos scan-chromosome.c:103-139: one CTA")), ("other", find("void launch_scan"))]
print("total samples", ts, "warp-instr", te)
for i, (name, start) in enumerate(marks):
    end = marks[i + 1][1] if i + 1 < len(marks) else 10 ** 9
    s = sum(a[0] for ln, a in agg.items() if start <= ln < end); e = sum(a[1] for ln, a in agg.items() if start <= ln < end)
    print(f"{name:40s} lines {start:4d}+  samples {100*s/ts:5.1f}%  inst {100*e/te:5.1f}%")
