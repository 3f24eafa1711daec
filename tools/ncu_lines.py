"""Per CUDA source line: stall samples and executed instructions from `ncu --page source --print-source cuda,sass --csv`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = None
agg = {}
for r in rows:
    if len(r) > 8 and r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) < 8: continue
    if r[2] != "-": continue          # SASS rows carry an address; the line rows have '-'
    try:
        ln = int(r[0]); samp = int(r[6] or 0); ex = int(r[7] or 0)
    except ValueError:
        continue
    a = agg.setdefault(ln, [r[1], 0, 0]); a[1] += samp; a[2] += ex
ts = sum(a[1] for a in agg.values()); te = sum(a[2] for a in agg.values())
print("total samples", ts, "warp-instr", te)
top = sorted(agg.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 30]
for ln, (src, s, e) in sorted(top):
    print(f"{ln:4d} samples {100*s/ts:5.1f}%  inst {100*e/te:5.1f}%  {src.strip()[:110]}")
