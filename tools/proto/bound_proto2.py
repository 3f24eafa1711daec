"""Prototype 2: prefix-max (Q) bounds for the candidates whose cut lies beyond a tile's last covering candidate."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from fscl_b200 import synth
from oracle import pyoracle as po
from moment_proto import logt
from bound_proto import piece_max

T = 2048

def main(scale=1.0, which="c2"):
    if which == "c2":
        tab, _ = synth.config2(scale=scale, planted=True)
        flat = po.RefProblem(tab).flatten()
    else:
        tab, _ = synth.config3(scale=scale)
        flat = po.RefProblem(tab, asc_depth=20, asc_min_freq=2).flatten()
    pos, cls = flat["pos"].astype(np.int64), flat["cls"]
    K, step = flat["n_knots"], flat["step"]
    coef = flat["coef"].copy()
    coef[:, :, 3] -= flat["null_tab"][:, None]
    n_cls = coef.shape[0]
    with np.errstate(all="ignore"):
        M = piece_max(coef, step, K, 1)
    Q = np.maximum.accumulate(M, axis=1)
    LT = po.log_table()
    las = [-20.0]
    for _ in range(10): las.append(las[-1] + 2.4)
    dmaxs = [np.exp(4.0 - la) for la in las]
    N = len(pos)
    rng = np.random.default_rng(1)
    lo_, hi_ = flat["chr_start"][0], flat["chr_start"][0] + flat["chr_nsnps"][0]
    pos, cls = pos[lo_:hi_], cls[lo_:hi_]; N = len(pos)
    sweeps = [int(pos[N // 2])] + list(rng.integers(pos[0], pos[-1], 12))
    for sweep in sweeps:
        near = int(np.argmin(np.abs(pos - sweep)))
        ws, we = max(near - 81920, 0), min(near + 81920, N - 1)
        d_all = np.abs(pos - sweep); x0_all = logt(LT, d_all)
        S = np.zeros(11); UB = np.zeros(11); exact_terms = np.zeros(11, dtype=np.int64); isexact = np.ones(11, dtype=bool)
        for j, la in enumerate(las):
            idx = np.arange(ws, we + 1); idx = idx[d_all[idx] <= dmaxs[j]]
            x = x0_all[idx] + la
            k = np.clip(((x + 20.0) / step).astype(np.int64), 0, K - 1)
            c = coef[cls[idx], k]
            y = x * (x * (x * c[:, 0] + c[:, 1]) + c[:, 2]) + c[:, 3]
            S[j] = y.sum()
            bounded = np.zeros(N, dtype=bool); ub = 0.0
            for t in range(ws // T, we // T + 1):
                lo, hi = t * T, min(t * T + T, N)
                if lo < ws or hi - 1 > we: continue
                dfar = d_all[lo:hi].max(); dnr = d_all[lo:hi].min()
                jf = sum(1 for jj in range(11) if dmaxs[jj] >= dfar)
                if j >= jf: continue
                cnt = np.bincount(cls[lo:hi], minlength=n_cls); nz = cnt > 0
                xb = logt(LT, np.array([dfar]))[0] + la; xa = logt(LT, np.array([dnr]))[0] + la
                kb = int(min(max((xb + 20.0) / step, 0), K - 1)); ka = int(min(max((xa + 20.0) / step, 0), K - 1))
                if j <= jf - 2 and j < 3:
                    ub += float((cnt[nz] * Q[nz, kb]).sum()); bounded[lo:hi] = True
                elif j == jf - 1 and j < 4 and kb - ka <= 3 and not (pos[lo] <= sweep <= pos[hi - 1]):
                    ub += float((cnt[nz] * M[nz, ka:kb + 1].max(axis=1)).sum()); bounded[lo:hi] = True
            isexact[j] = not bounded.any()
            ub += float(y[~bounded[idx]].sum()); exact_terms[j] = int((~bounded[idx]).sum())
            UB[j] = ub
        best = S[isexact].max()
        msg = " ".join(f"[{j}] S {S[j]:.4g} UB {UB[j]:.4g}" for j in range(11) if not isexact[j])
        print(f"sweep {sweep:>9d} best {best:.5g} exact terms {exact_terms.sum()} of {sum((d_all[ws:we+1] <= dm).sum() for dm in dmaxs)}  {msg}  ok={all(UB[j] < best for j in range(11) if not isexact[j])}")

if __name__ == "__main__":
    main(float(sys.argv[1]) if len(sys.argv) > 1 else 1.0, sys.argv[2] if len(sys.argv) > 2 else "c2")
