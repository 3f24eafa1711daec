"""Per-kernel launches / total time / share from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import csv, re, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("==")) ]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
H = rows[hdr]
ik, iv, iu = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
agg = {}
for r in rows[hdr + 1:]:
    if len(r) <= iv: continue
    v = float(r[iv].replace(",", ""))
    ms = v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "msecond": 1.0, "ms": 1.0, "second": 1e3}.get(r[iu], 1e-6)
    name = re.sub(r"\(.*", "", r[ik]).strip()
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += ms
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':60s} {'launches':>9s} {'ms':>12s} {'share':>8s}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:60s} {a[0]:9d} {a[1]:12.3f} {a[1]/tot:8.4f}")
