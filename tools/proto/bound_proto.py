"""Prototype: upper bounds of the far-tile sums of an alpha candidate from per-tile class counts and
per-(class, knot) maxima of the spline pieces.  How tight are they against the best exact candidate?"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from fscl_b200 import synth
from oracle import pyoracle as po
from moment_proto import logt

T = 2048

def piece_max(coef, step, K, sub=1):
    """max over each knot interval (optionally `sub` sub-intervals) of the cubic piece: end points + interior critical points"""
    n_cls = coef.shape[0]
    M = np.empty((n_cls, K * sub))
    for k in range(K):
        c0, c1, c2, c3 = (coef[:, k, i] for i in range(4))
        for s in range(sub):
            a = -20.0 + (k + s / sub) * step
            b = -20.0 + (k + (s + 1) / sub) * step
            if k == K - 1 and s == sub - 1: b = 4.0 + 1e-9
            f = lambda x: x * (x * (x * c0 + c1) + c2) + c3
            m = np.maximum(f(a), f(b))
            # critical points of 3c0 x^2 + 2c1 x + c2
            disc = 4 * c1 * c1 - 12 * c0 * c2
            with np.errstate(all="ignore"):
                sq = np.sqrt(np.maximum(disc, 0))
                for sgn in (1, -1):
                    r = (-2 * c1 + sgn * sq) / (6 * c0)
                    ok = (disc >= 0) & (r > a) & (r < b) & np.isfinite(r)
                    m = np.where(ok, np.maximum(m, f(np.where(ok, r, a))), m)
            M[:, k * sub + s] = m
    return M

def main(scale=1.0, sub=1):
    tab, _ = synth.config2(scale=scale, planted=True)
    flat = po.RefProblem(tab).flatten()
    pos, cls = flat["pos"].astype(np.int64), flat["cls"]
    K, step = flat["n_knots"], flat["step"]
    coef = flat["coef"].copy()
    coef[:, :, 3] -= flat["null_tab"][:, None]
    n_cls = coef.shape[0]
    M = piece_max(coef, step, K, sub)
    LT = po.log_table()
    inv = sub / step
    las = [-20.0]
    for _ in range(10): las.append(las[-1] + 2.4)
    N = len(pos)
    rng = np.random.default_rng(1)
    sweeps = [int(50_000_000 * scale), int(50_050_000 * scale)] + list(rng.integers(pos[0], pos[-1], 10))
    for sweep in sweeps:
        near = int(np.argmin(np.abs(pos - sweep)))
        ws, we = max(near - 81920, 0), min(near + 81920, N - 1)
        if ws == 0: we = min(2 * 81920, N - 1)
        if we == N - 1: ws = max(N - 1 - 2 * 81920, 0)
        d_all = np.abs(pos - sweep)
        x0_all = logt(LT, d_all)
        S, UB, nb = [], [], []
        for j, la in enumerate(las):
            idx = np.arange(ws, we + 1)
            idx = idx[x0_all[idx] + la <= 4.0]
            if d_all[near] > 0 and x0_all[near] + la > 4.0: idx = idx[:0]
            x = x0_all[idx] + la
            k = np.clip(((x + 20.0) / step).astype(np.int64), 0, K - 1)
            c = coef[cls[idx], k]
            y = x * (x * (x * c[:, 0] + c[:, 1]) + c[:, 2]) + c[:, 3]
            S.append(float(y.sum()))
            # bound: far tiles (inside window, one side, wholly within the cut) bounded, the rest exact
            ub = 0.0; bounded = np.zeros(N, dtype=bool); ntb = 0
            dmax = np.exp(4.0 - la)
            for t in range(ws // T, we // T + 1):
                lo, hi = t * T, min(t * T + T, N)
                if lo < ws or hi - 1 > we: continue
                if not (pos[hi - 1] < sweep or pos[lo] > sweep): continue
                xs = x0_all[[lo, hi - 1]] + la
                if xs.max() > 4.0: continue
                klo = int(min(max((xs.min() + 20.0) * inv, 0), K * sub - 1)); khi = int(min(max((xs.max() + 20.0) * inv, 0), K * sub - 1))
                if khi - klo > 3 * sub: continue       # wide tiles stay exact
                cnt = np.bincount(cls[lo:hi], minlength=n_cls)
                nz = cnt > 0; ub += float((cnt[nz] * M[nz, klo:khi + 1].max(axis=1)).sum())
                bounded[lo:hi] = True; ntb += 1
            ub += float(y[~bounded[idx]].sum())
            UB.append(ub); nb.append(ntb)
        S, UB = np.array(S), np.array(UB)
        exact = np.array(nb) == 0
        best = S[exact].max() if exact.any() else np.inf
        msg = " ".join(f"[{j}] S {S[j]:.4g} UB {UB[j]:.4g} tiles {nb[j]}" for j in range(11) if not exact[j])
        print(f"sweep {sweep:>9d} best_exact {best:.6g} (argmax all {int(np.argmax(S))})  {msg}  pruned: {[j for j in range(11) if not exact[j] and UB[j] < best]}")

if __name__ == "__main__":
    main(float(sys.argv[1]) if len(sys.argv) > 1 else 1.0, int(sys.argv[2]) if len(sys.argv) > 2 else 1)
