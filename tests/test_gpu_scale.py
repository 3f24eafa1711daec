"""GPU tests at BASELINE.json sizes and shapes: full-size c2 through size-independent
properties plus oracle spot checks; c4-shaped (22 chromosomes) and c5-shaped
(11 depths, folded mix, >1500 classes) inputs against the oracle."""
import ctypes as C

import numpy as np
import pytest

from fscl_b200 import synth
from oracle import pyoracle as po
from tests.common import clr_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api(built):
    from fscl_b200 import api as a
    return a


@pytest.fixture(scope="module")
def c2(api):
    tab, opt = synth.config2()
    flat = po.RefProblem(tab).flatten()
    return flat, api.Scanner(flat), opt["G"]


def test_c2_full_size_scan_spot_checked_and_deterministic(api, c2):
    flat, sc, G = c2
    a = sc.scan_chromosome(G)
    b = sc.scan_chromosome(G)
    assert len(a) == 1000 and a.tobytes() == b.tobytes()                 # run-to-run identical
    assert (a["clr"] >= 0).all() and (np.diff(a["sweep_pos"]) >= 0).all()
    assert ((a["n_snps"] == a["window_end"] - a["window_start"] + 1) & (a["n_snps"] <= 2 * 81920 + 1)).all()
    o = po.Oracle(flat, mode=2)
    cells = o.enum_cells(G)
    for i in (0, 137, 499, 500, 862, 999):                              # ends, middle, planted sweep
        w, ne = o.search_maxpos(int(cells[0, i]), int(cells[1, i]), int(cells[2, i]))
        assert ne == 12
        assert (a[i]["sweep_pos"], a[i]["window_start"], a[i]["window_end"]) == (w["sweep_pos"], w["wstart"], w["wend"])
        assert a[i]["lalpha"] == w["lalpha"] and clr_close(a[i]["clr"], w["clr"])


def test_c2_full_size_block_permutation_properties(api, c2):
    flat, sc, G = c2
    o = po.Oracle(flat)
    g = o.rng("philox", 2024)
    n = len(flat["pos"])
    for t in (0, 5):
        got, blocks = sc.block_permute(t, rng=api.RNG_PHILOX, seed=2024)
        assert np.array_equal(np.sort(got), np.sort(flat["cls"]))        # a permutation of the classes
        assert blocks[:, 2].sum() == n and blocks[0, 0] == 0             # i-ranges tile [0, n)
        assert (blocks[:, 0] + blocks[:, 2] <= n).all() and (blocks[:, 1] + blocks[:, 2] <= n).all()
        assert (blocks[:, 1] >= 0).all()
        o.L.fo_rng_set_perm(C.addressof(g), t)
        want, _, _ = o.block_permute(g)
        assert np.array_equal(got, want)                                 # bit-exact vs the restatement at full size
        dev, flags = sc.device_gen_blocks(t, seed=2024)
        assert np.array_equal(dev, blocks) and not (flags & 6)


def test_c2_full_size_permuted_scans_invariant_to_round_size(api, c2):
    flat, sc, G = c2
    obs = sc.scan_chromosome(G)
    a, sta, _ = sc.scan_permute(obs, 7, G, rng=api.RNG_PHILOX, seed=8, prune=0, round_perms=1)
    b, stb, _ = sc.scan_permute(obs, 7, G, rng=api.RNG_PHILOX, seed=8, prune=0, round_perms=8)
    assert (a["permute_n"] == 8).all() and np.array_equal(a["permute_p"], b["permute_p"])
    # exceedance counts of a neutral-ish genome: p-values roughly uniform, planted sweep cell extreme
    assert a["permute_p"][obs["clr"].argmax()] == 0


@pytest.mark.parametrize("shape", ["c4", "c5"])
def test_genome_shapes_match_oracle(api, shape):
    if shape == "c4":
        tab, opt = synth.config4(scale=0.01)          # 22 chromosomes, 30k sites
    else:
        tab, opt = synth.config5(scale=0.002)         # 22 chromosomes, 40k sites, depths 80..100, 5 % folded
    flat = po.RefProblem(tab).flatten()
    G = opt["G"] * (4 if shape == "c5" else 1)
    sc = api.Scanner(flat)
    o = po.Oracle(flat, mode=2)
    got = sc.scan_chromosome(G)
    want, _ = o.scan(G)
    assert len(got) == len(want) and len(np.unique(got["chr"])) == 22
    for a, b in (("chr", "chr"), ("sweep_pos", "sweep_pos"), ("window_start", "wstart"), ("window_end", "wend"),
                 ("lalpha", "lalpha")):
        assert np.array_equal(got[a], want[b]), a
    assert clr_close(got["clr"], want["clr"]).all()
    if shape == "c5":
        assert len(flat["null_tab"]) > 1500 and len(flat["depths"]) == 11
    n_perm = 22
    w2, scored, flags, _, _ = o.scan_permute(want, n_perm, G, o.rng("philox", 31))
    g2, st, _ = sc.scan_permute(got, n_perm, G, rng=api.RNG_PHILOX, seed=31)
    assert np.array_equal(g2["permute_p"], w2["permute_p"]) and np.array_equal(g2["permute_n"], w2["permute_n"])
    assert np.array_equal(g2["permute_finished"], w2["finished"]) and st["n_scored"] == scored


def _same_points(a, b):
    return a.tobytes() == b.tobytes()


def test_c2_candidate_pruning_changes_nothing(api, c2):
    """Pruning dominated coarse candidates by their upper bounds must leave every
    output bit unchanged: forward scan, permuted scans, counts; and it must actually prune."""
    flat, sc, G = c2
    try:
        sc.set_exact_candidates(True)
        a = sc.scan_chromosome(G)
        terms_exact, _ = sc.bench_scan(G, 0, count_terms=True)
        pa, _, _ = sc.scan_permute(a, 5, G, rng=api.RNG_PHILOX, seed=77, prune=0)
        sc.set_exact_candidates(False)
        sc.reset_stats()
        b = sc.scan_chromosome(G)
        failures = sc.stats()["bound_failures"]
        terms_pruned, _ = sc.bench_scan(G, 0, count_terms=True)
        pb, _, _ = sc.scan_permute(b, 5, G, rng=api.RNG_PHILOX, seed=77, prune=0)
    finally:
        sc.set_exact_candidates(False)
    assert _same_points(a, b) and _same_points(pa, pb)
    assert 0 < terms_pruned < 0.5 * terms_exact          # the far tiles of the three widest candidates dominate c2
    # at the planted sweep a bounded candidate wins, so the exact pass has to run there -- and only there
    assert 0 < failures < 0.05 * 12 * len(b)


@pytest.mark.parametrize("shape", ["mid100", "c4", "c5", "sweep"])
def test_candidate_pruning_changes_nothing_on_other_shapes(api, shape):
    if shape == "mid100":
        tab, opt = synth.config2(scale=0.1)
        G = opt["G"]
    elif shape == "c4":
        tab, opt = synth.config4(scale=0.02)
        G = opt["G"]
    elif shape == "c5":
        tab, opt = synth.config5(scale=0.004)
        G = opt["G"] * 4
    else:
        # a sweep footprint tens of Mb wide: wide (low-alpha) candidates score high, so the bounds
        # cannot exclude them everywhere and the exact pass has to run
        tab, opt = synth.config2(scale=0.5, planted=False)
        rng = np.random.default_rng(3)
        mid = int(tab["pos"][len(tab["pos"]) // 2])
        hit = (np.abs(tab["pos"] - mid) < 12_000_000) & (rng.random(len(tab["count"])) < 0.7)
        tab["count"][hit] = np.where(rng.random(int(hit.sum())) < 0.5, rng.integers(1, 3, int(hit.sum())),
                                     rng.integers(97, 100, int(hit.sum())))
        G = opt["G"]
    flat = po.RefProblem(tab).flatten()
    sc = api.Scanner(flat)
    sc.set_exact_candidates(True)
    a = sc.scan_chromosome(G)
    pa, _, _ = sc.scan_permute(a, 3, G, rng=api.RNG_PHILOX, seed=5, prune=0)
    sc.set_exact_candidates(False)
    sc.reset_stats()
    b = sc.scan_chromosome(G)
    failures = sc.stats()["bound_failures"]
    pb, _, _ = sc.scan_permute(b, 3, G, rng=api.RNG_PHILOX, seed=5, prune=0)
    assert _same_points(a, b) and _same_points(pa, pb)
    if shape == "sweep":
        assert failures > 0                               # bounds that could not exclude their candidate
        o = po.Oracle(flat, mode=2)
        want, _ = o.scan(G)
        assert np.array_equal(b["sweep_pos"], want["sweep_pos"]) and np.array_equal(b["lalpha"], want["lalpha"])
        assert clr_close(b["clr"], want["clr"]).all()


@pytest.mark.parametrize("shape", ["c2", "c4"])
def test_endpoint_sharing_changes_nothing(api, c2, shape):
    """A cell that starts where the previous one ended takes that end point's evaluation over (scan_cells_kernel);
    switched off, every cell evaluates both end points itself as search_maxpos does (scan-chromosome.c:103-107).
    Every output bit must agree: forward scan, stored permuted CLRs, counts -- with and without a pruned test
    (late rounds hold isolated cells, early ones whole chromosomes) -- and the shared form must evaluate fewer terms."""
    if shape == "c2":
        flat, sc, G = c2
    else:
        tab, opt = synth.config4(scale=0.05)
        flat, G = po.RefProblem(tab).flatten(), opt["G"]
        sc = api.Scanner(flat)
    try:
        sc.set_endpoint_sharing(False)
        a = sc.scan_chromosome(G)
        ta, _ = sc.bench_scan(G, 16, count_terms=True)
        pa, _, ca = sc.scan_permute(a, 24, G, rng=api.RNG_PHILOX, seed=41, prune=0, save_clr=True)
        qa, _, da = sc.scan_permute(a, 60, G, rng=api.RNG_PHILOX, seed=43, prune=1, save_clr=True)
        sc.set_endpoint_sharing(True)
        b = sc.scan_chromosome(G)
        tb, _ = sc.bench_scan(G, 16, count_terms=True)
        pb, _, cb = sc.scan_permute(b, 24, G, rng=api.RNG_PHILOX, seed=41, prune=0, save_clr=True)
        qb, _, db = sc.scan_permute(b, 60, G, rng=api.RNG_PHILOX, seed=43, prune=1, save_clr=True)
    finally:
        sc.set_endpoint_sharing(True)
    assert _same_points(a, b) and _same_points(pa, pb) and _same_points(qa, qb)
    assert ca.tobytes() == cb.tobytes() and da.tobytes() == db.tobytes()
    assert (qa["permute_finished"] == 1).any() and (qa["permute_finished"] == 0).any()   # sparse late rounds were run
    # launches of more than 8192 tasks (16 genomes here, the rounds of 17+ permutations above) hold chains of up to 8
    # cells; the last 8192 tasks of a launch run in shorter chains and singly (scan_cells_kernel)
    assert 0.85 * ta < tb < 0.99 * ta
