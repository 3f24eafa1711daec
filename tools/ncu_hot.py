"""Hot SASS regions of an ncu source-page CSV: groups of consecutive instructions with equal execution count."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, isamp, iexec = hdr.index('Address'), hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
data = []
for r in rows[2:]:
    try: data.append((r[ia], r[isrc], int(r[isamp] or 0), int(r[iexec] or 0)))
    except Exception: pass
tot_s = sum(d[2] for d in data); tot_e = sum(d[3] for d in data)
print('total samples', tot_s, 'total inst', tot_e, 'n', len(data))
groups = []
for d in data:
    if groups and groups[-1][1] == d[3]: groups[-1][2] += 1; groups[-1][3] += d[2]; groups[-1][4].append(d[1])
    else: groups.append([d[0], d[3], 1, d[2], [d[1]]])
for g in sorted(groups, key=lambda g: -g[3])[:int(sys.argv[2]) if len(sys.argv) > 2 else 16]:
    print(f"addr {g[0][-5:]} exec/instr {g[1]:>12d} n_instr {g[2]:>4d} inst {g[1]*g[2]/tot_e*100:5.1f}% samples {g[3]/tot_s*100:5.1f}%  first: {g[4][0][:50]}")
if len(sys.argv) > 3:
    start = [i for i, d in enumerate(data) if d[0].endswith(sys.argv[3])][0]
    for d in data[start:start + int(sys.argv[4])]:
        print(d[0][-5:], f"{d[3]:>10d} {d[2]:>7d}  {d[1][:90]}")
