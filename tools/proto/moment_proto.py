"""Numerical prototype (CPU, numpy) of the moment evaluation of far tiles.

For one sweep position of a c2-shaped problem it compares, per coarse alpha,
 (a) long-double sum of per-site double terms  (oracle mode 1 arithmetic)
 (b) double sum of per-site terms              (per-site GPU path)
 (c) far tiles by (class, knot-interval) cells: centred moments + Taylor re-expansion
against each other and against the reference accumulator's quantum ulp(N).
"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from fscl_b200 import synth
from oracle import pyoracle as po

T = 2048
SPAN = 16

def logt(Tab, d):
    d = np.asarray(d)
    sh = np.where(d > 0xFFFFFF, 16, np.where(d > 0xFFFF, 8, 0))
    c = np.where(d > 0xFFFFFF, 11.783502069519070, np.where(d > 0xFFFF, 5.545177444479562, 0.0))
    t = Tab[d >> sh]
    return np.where(sh > 0, c + t, t)

def main(scale=1.0, sweeps=(50_000_000, 20_000_123, 77_000_777)):
    tab, _ = synth.config2(scale=scale, planted=False)
    flat = po.RefProblem(tab).flatten()
    pos, cls = flat["pos"].astype(np.int64), flat["cls"]
    K, step = flat["n_knots"], flat["step"]
    coef = flat["coef"].copy()
    coef[:, :, 3] -= flat["null_tab"][:, None]
    LT = po.log_table()
    inv_step = 1.0 / step
    las = [-20.0]
    for _ in range(10): las.append(las[-1] + 2.4)
    N = len(pos)
    for sweep in sweeps:
        sweep = int(sweep * scale)
        near = int(np.argmin(np.abs(pos - sweep)))
        ws, we = max(near - 81920, 0), min(near + 81920, N - 1)
        Nnull = float(flat["null_tab"][cls[ws:we + 1]].sum())
        quantum = np.spacing(abs(Nnull))
        print(f"sweep {sweep} window {ws}..{we} null {Nnull:.1f} quantum {quantum:.3e}")
        d_all = np.abs(pos - sweep)
        x0_all = logt(LT, d_all)
        for j, la in enumerate(las[:4]):
            dmax = np.exp(4.0 - la)
            # per-site terms over the window (only alphas whose cut covers the window are interesting here)
            idx = np.arange(ws, we + 1)
            idx = idx[x0_all[idx] + la <= 4.0]
            x = x0_all[idx] + la
            k = np.clip(((x + 20.0) * inv_step).astype(np.int64), 0, K - 1)
            c = coef[cls[idx], k]
            y = x * (x * (x * c[:, 0] + c[:, 1]) + c[:, 2]) + c[:, 3]
            a = float(np.sum(y.astype(np.longdouble)))
            b = float(np.sum(y))
            mag = np.abs(c[:, 3]) + np.abs(c[:, 2] * x) + np.abs(c[:, 1] * x * x) + np.abs(c[:, 0] * x ** 3)
            # (c) moments on far tiles
            tot = np.longdouble(0.0)
            n_cells = n_far_sites = 0
            t0, t1 = ws // T, we // T
            done = np.zeros(N, dtype=bool)
            for t in range(t0, t1 + 1):
                lo, hi = t * T, min(t * T + T, N)
                if lo < ws or hi - 1 > we: continue
                if not (pos[hi - 1] < sweep or pos[lo] > sweep): continue
                if d_all[lo:hi].max() > dmax: continue
                if x0_all[lo:hi].max() + la > 4.0: continue
                # far tile: sort by (class, index), spans of SPAN, cells = (class, knot) runs within spans
                order = np.lexsort((np.arange(lo, hi), cls[lo:hi])) + lo
                xr = 0.5 * (x0_all[lo:hi].min() + x0_all[lo:hi].max())
                X = xr + la
                u = x0_all[order] - xr
                kk = np.clip(((X + u + 20.0) * inv_step).astype(np.int64), 0, K - 1)
                cc = cls[order]
                span = np.arange(hi - lo) // SPAN
                key_change = np.ones(hi - lo, dtype=bool)
                key_change[1:] = (cc[1:] != cc[:-1]) | (kk[1:] != kk[:-1]) | (span[1:] != span[:-1])
                starts = np.nonzero(key_change)[0]
                ends = np.append(starts[1:], hi - lo)
                tile_sum = 0.0
                for s, e in zip(starts, ends):
                    uu = u[s:e]
                    U1 = 0.0; U2 = 0.0; U3 = 0.0
                    for v in uu:            # sequential double accumulation like a thread would
                        U1 += v; v2 = v * v; U2 += v2; U3 += v2 * v
                    c0, c1, c2, c3 = coef[cc[s], kk[s]]
                    t3 = 3.0 * c0 * X
                    p = X * (X * (X * c0 + c1) + c2) + c3
                    p1 = (t3 + 2.0 * c1) * X + c2
                    p2 = t3 + c1
                    tile_sum += (e - s) * p + p1 * U1 + p2 * U2 + c0 * U3
                n_cells += len(starts); n_far_sites += hi - lo
                tot += np.longdouble(tile_sum)
                done[lo:hi] = True
            rest = ~done[idx]
            tot += np.sum(y[rest].astype(np.longdouble))
            cm = float(tot)
            print(f"  la {la:6.1f} terms {len(idx):6d} S {a:+.12e}  b-a {b - a:+.2e}  moments-a {cm - a:+.2e}"
                  f"  far {n_far_sites} cells {n_cells} ({n_far_sites / max(n_cells, 1):.1f}/cell)"
                  f"  term mag mean {mag.mean():.2e} max {mag.max():.2e}")

if __name__ == "__main__":
    main(float(sys.argv[1]) if len(sys.argv) > 1 else 1.0)
