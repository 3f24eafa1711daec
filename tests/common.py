"""Shared problem builders for the tests (inputs are prepared by the compiled,
unmodified reference: oracle/_ref/libfscl_ref.so, which travels to the GPU box)."""
import functools

import numpy as np

from fscl_b200 import synth
from oracle import pyoracle as po

CLR_RTOL = 1e-9      # north_star: 1e-9 relative for FP64; stated as |d| <= 1e-9 * max(|CLR|, 1)


def clr_close(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    with np.errstate(invalid="ignore"):
        ok = np.abs(a - b) <= CLR_RTOL * np.maximum(np.maximum(np.abs(a), np.abs(b)), 1.0)
    return ok | (a == b) | (np.isnan(a) & np.isnan(b))


@functools.lru_cache(maxsize=None)
def problem(kind):
    """name -> (table, flat, G)"""
    if kind == "small50":          # c1-like: n=50, unfolded, one chromosome
        tab = synth.single_chromosome(4000, 2_000_000, 50, 11, planted=True)
        return tab, po.RefProblem(tab).flatten(), 100_000
    if kind == "c1":               # config 1: ms-like 1 Mb, n=50, G=1 kb
        tab, opt = synth.config1()
        return tab, po.RefProblem(tab).flatten(), opt["G"]
    if kind == "mid100":           # c2 shape at 1/10 scale
        tab, opt = synth.config2(scale=0.1)
        return tab, po.RefProblem(tab).flatten(), opt["G"]
    if kind == "asc":              # c3 shape: 2 chr, 5 depths, folded mix, K=2 of M=20
        tab, opt = synth.config3(scale=0.05)
        return tab, po.RefProblem(tab, asc_depth=20, asc_min_freq=2).flatten(), opt["G"]
    if kind == "ascnan":           # depths for which the reference's K=2 tables are NaN (D14)
        tab, opt = synth.config3(scale=0.05, depths=(100, 98, 96, 94, 90))
        return tab, po.RefProblem(tab, asc_depth=20, asc_min_freq=2).flatten(), opt["G"]
    if kind == "asc1":             # K=1 of M=20
        tab, opt = synth.config3(scale=0.05, depths=(100, 98, 96, 94, 90))
        return tab, po.RefProblem(tab, asc_depth=20, asc_min_freq=1).flatten(), opt["G"]
    if kind == "invariant":        # --include-invariant: sites with count 0 / depth kept, their classes in the tables
        rng = np.random.default_rng(9)
        tab = synth.single_chromosome(6000, 3_000_000, 30, 12, planted=True)
        inv = rng.random(6000) < 0.15
        tab["count"][inv] = np.where(rng.random(int(inv.sum())) < 0.5, 0, 30)
        return tab, po.RefProblem(tab, include_invariant=1).flatten(), 100_000
    if kind == "folded":           # every site folded (minor allele counts), two depths
        rng = np.random.default_rng(10)
        tab = synth.single_chromosome(8000, 4_000_000, 40, 13, planted=True)
        tab["depth"] = tab["depth"].copy()
        tab["depth"][::3] = 36
        tab["count"] = np.minimum(tab["count"], tab["depth"] - tab["count"]).astype(np.int32)
        tab["count"] = np.maximum(tab["count"], 1).astype(np.int32)
        tab["folded"] = np.ones(len(tab["pos"]), dtype=np.int32)      # (chr and folded share one zero array in synth)
        return tab, po.RefProblem(tab).flatten(), 100_000
    if kind == "genome":           # c4 shape: 22 chromosomes, tiny scale
        tab, opt = synth.config4(scale=0.002)
        return tab, po.RefProblem(tab).flatten(), opt["G"]
    if kind == "dense":            # coincident grid positions (D3) + a 1-SNP chromosome
        rng = np.random.default_rng(5)
        pos1 = np.arange(1000, 1000 + 3000 * 40, 40)            # every grid start hits a SNP
        pos1 = np.unique(np.concatenate([pos1, pos1[::7] + 1, pos1[::7] + 2]))
        pos2 = np.array([5000])
        pos3 = np.sort(rng.choice(np.arange(1, 400_000), 1500, replace=False))
        pos = np.concatenate([pos1, pos2, pos3])
        chr_of = np.concatenate([np.zeros(len(pos1)), np.ones(1), np.full(len(pos3), 2)]).astype(np.int32)
        depth = np.full(len(pos), 30)
        cnt = synth._neutral_counts(rng, len(pos), depth)
        tab = synth.make_table(["cA", "cB", "cC"], chr_of, pos, cnt, depth, np.zeros(len(pos), dtype=np.int32))
        return tab, po.RefProblem(tab).flatten(), 20_000
    raise KeyError(kind)
