"""Synthetic SNP-table generators for the five BASELINE.json configs.

Inputs follow SURVEY.md §8d: 5-column text ``chr pos count sample_size folded``
(the format parsed at /root/reference/snp-input.c:49-66), unique sorted
positions, no header line.  Everything is drawn from
``numpy.random.Generator(PCG64(seed))`` so a (config, seed, scale) triple is
reproducible on any box.
"""
from __future__ import annotations

import numpy as np

HG19_MB = [249, 243, 198, 191, 181, 171, 159, 146, 141, 136, 135,
           134, 115, 107, 103, 90, 81, 78, 59, 63, 48, 51]


def _neutral_counts(rng, n_sites, depth):
    """Derived-allele counts with P(f) ∝ 1/f on 1..depth-1 (vector of depths ok)."""
    depth = np.broadcast_to(np.asarray(depth), (n_sites,))
    out = np.empty(n_sites, dtype=np.int64)
    for d in np.unique(depth):
        sel = np.nonzero(depth == d)[0]
        w = 1.0 / np.arange(1, d)
        out[sel] = rng.choice(np.arange(1, d), size=sel.size, p=w / w.sum())
    return out


def _positions(rng, n_sites, length_bp):
    """Sorted unique positions in [1, length_bp]."""
    pos = np.unique(rng.integers(1, length_bp + 1, size=int(n_sites * 1.02) + 16))
    while pos.size < n_sites:
        pos = np.unique(np.concatenate([pos, rng.integers(1, length_bp + 1, size=n_sites)]))
    pos = rng.permutation(pos)[:n_sites]
    pos.sort()
    return pos.astype(np.int64)


def make_table(chr_names, chr_of, pos, count, depth, folded):
    return dict(chr_names=list(chr_names), chr=np.asarray(chr_of, dtype=np.int32),
                pos=np.asarray(pos, dtype=np.int32), count=np.asarray(count, dtype=np.int32),
                depth=np.asarray(depth, dtype=np.int32), folded=np.asarray(folded, dtype=np.int32))


def single_chromosome(n_sites, length_bp, depth, seed, planted=False, name="chr1"):
    """c1/c2 shape: one chromosome, constant depth, unfolded, neutral 1/f."""
    rng = np.random.Generator(np.random.PCG64(seed))
    pos = _positions(rng, n_sites, length_bp)
    cnt = _neutral_counts(rng, n_sites, depth)
    if planted:
        mid, half = length_bp // 2, max(length_bp // 500, 2000)
        near = np.nonzero(np.abs(pos - mid) < half)[0]
        hit = near[rng.random(near.size) < 0.6]
        u = rng.random(hit.size)
        cnt[hit] = np.where(u < 0.5, rng.integers(1, 4, hit.size),
                            rng.integers(depth - 3, depth, hit.size))
    z = np.zeros(n_sites, dtype=np.int32)
    return make_table([name], z, pos, cnt, np.full(n_sites, depth), z)


def config1(seed=101, scale=1.0):
    """ms-like neutral 1 Mb, n=50, ~1800 sites, -G 1000."""
    n_sites = max(int(1800 * scale), 50)
    return single_chromosome(n_sites, int(1_000_000 * scale) or 1000, 50, seed), dict(G=1000)


def config2(seed=102, scale=1.0, planted=True):
    """200k SNPs / 100 Mb / n=100, -G 100000 -p 1000."""
    n_sites = max(int(200_000 * scale), 100)
    return single_chromosome(n_sites, max(int(100_000_000 * scale), 10_000), 100, seed,
                             planted=planted), dict(G=100_000, p=1000)


def config3(seed=103, scale=1.0, n_chr=2, depths=(98, 88, 86, 82, 78)):
    """Array-style: variable depth, K=2 of M=20 ascertainment, 30 % folded.

    The default depths are ones for which the reference's ascertainment
    probability of a singleton (asc-bias.c:12-25: 1 - C(n,20)/C(n,20), pure
    rounding noise) comes out positive; for depths where it comes out negative
    (e.g. 100, 96, 94, 90) the reference's tables are NaN."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n_sites = max(int(100_000 * scale), 200)
    length = max(int(100_000_000 * scale), 20_000)
    per = [n_sites // n_chr] * n_chr
    per[-1] += n_sites - sum(per)
    chr_of, pos = [], []
    for c, m in enumerate(per):
        pos.append(_positions(rng, m, length // n_chr))
        chr_of.append(np.full(m, c))
    chr_of, pos = np.concatenate(chr_of), np.concatenate(pos)
    depth = rng.choice(list(depths), size=n_sites, p=[0.6, 0.1, 0.1, 0.1, 0.1])
    cnt = _neutral_counts(rng, n_sites, depth)
    bad = (cnt < 2) | (cnt > depth - 2)
    while bad.any():
        cnt[bad] = _neutral_counts(rng, int(bad.sum()), depth[bad])
        bad = (cnt < 2) | (cnt > depth - 2)
    folded = (rng.random(n_sites) < 0.3).astype(np.int32)
    cnt = np.where(folded == 1, np.minimum(cnt, depth - cnt), cnt)
    names = [f"chr{c + 1}" for c in range(n_chr)]
    return make_table(names, chr_of, pos, cnt, depth, folded), dict(G=100_000, asc_depth=20, asc_min_freq=2)


def genome(n_sites, seed, depth_lo=100, depth_hi=100, frac_folded=0.0, scale=1.0):
    """c4/c5 shape: 22 autosomes with hg19-like lengths, sites ∝ length."""
    rng = np.random.Generator(np.random.PCG64(seed))
    lens = np.array([max(int(m * 1_000_000 * scale), 50_000) for m in HG19_MB], dtype=np.int64)
    n_sites = int(n_sites * scale)
    per = np.maximum((n_sites * lens / lens.sum()).astype(np.int64), 16)
    chr_of, pos = [], []
    for c, m in enumerate(per):
        pos.append(_positions(rng, int(m), int(lens[c])))
        chr_of.append(np.full(int(m), c))
    chr_of, pos = np.concatenate(chr_of), np.concatenate(pos)
    n = pos.size
    if depth_lo == depth_hi:
        depth = np.full(n, depth_hi)
    else:
        depth = rng.integers(depth_lo // 2, depth_hi // 2 + 1, size=n) * 2
    cnt = _neutral_counts(rng, n, depth)
    folded = (rng.random(n) < frac_folded).astype(np.int32)
    cnt = np.where(folded == 1, np.minimum(cnt, depth - cnt), cnt)
    names = [f"chr{c + 1}" for c in range(22)]
    return make_table(names, chr_of, pos, cnt, depth, folded)


def config4(seed=104, scale=1.0):
    return genome(3_000_000, seed, scale=scale), dict(G=100_000, p=10_000)


def config5(seed=105, scale=1.0):
    return genome(20_000_000, seed, depth_lo=80, depth_hi=100, frac_folded=0.05, scale=scale), \
        dict(G=10_000, p=100_000)


def write_snp_table(table, path):
    """Write the 5-column text the reference parses (snp-input.c:49-66)."""
    names = table["chr_names"]
    with open(path, "w") as fh:
        c, p, k, d, f = (table[x] for x in ("chr", "pos", "count", "depth", "folded"))
        lines = [f"{names[c[i]]}\t{p[i]}\t{k[i]}\t{d[i]}\t{f[i]}\n" for i in range(len(p))]
        fh.writelines(lines)
