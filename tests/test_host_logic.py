"""Host side of the product (libfscl_b200.so entry points that need no device)
against the oracle, the C-ABI surface, and the sharded permutation protocol
over a 2-rank gloo group.  CPU only."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from oracle import pyoracle as po
from tests.common import problem
from tests.test_golden import load_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def api(built):
    from fscl_b200 import api as a
    a.load_library()
    return a


def test_library_exports_every_declared_symbol(api):
    hdr = open(os.path.join(ROOT, "include", "fscl_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(fscl_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 20
    L = api.load_library()
    for name in declared:
        assert hasattr(L, name), name
    assert set(api.EXPORTS) == set(declared)
    assert C.sizeof(api.fscl_problem_t) == 128 and api.PT_DTYPE.itemsize == 72


def test_no_device_is_an_error_not_a_fallback(api):
    L = api.load_library()
    if L.fscl_device_count() > 0:
        pytest.skip("a GPU is present")
    _, flat, _ = problem("small50")
    with pytest.raises(api.FsclError, match="no CUDA device"):
        api.Scanner(flat)
    with pytest.raises(api.FsclError):
        api.fp64_peak_tflops()


def test_rng_known_answers(api):
    assert list(api.host_rng_draws(api.RNG_GLIBC, 0xFD821A6, 8)) == \
        [955590808, 761095611, 1240286151, 695883595, 1676369324, 1254833338, 1124954057, 1612196038]
    L = po.lib()
    for seed, t, stream in ((1, 0, 0), (0xFD821A6, 7, 1), (2**40 + 3, 123456, 0)):
        got = api.host_rng_draws(api.RNG_PHILOX, seed, 50, t=t, stream=stream)
        want = [L.fo_philox_draw(seed & 0xFFFFFFFF, seed >> 32, t, stream, i) for i in range(50)]
        assert list(got) == want


def test_alpha_tables_match_oracle(api):
    la, dmax, n = api.host_alpha_tables(po.log_table())
    L = po.lib()
    coarse, fine = np.zeros(16), np.zeros(16)
    assert L.fo_alpha_lists(coarse.ctypes.data, -1, None) == n[11] == 11
    assert np.array_equal(la[11, :11], coarse[:11])
    for m in range(11):
        k = L.fo_alpha_lists(coarse.ctypes.data, m, fine.ctypes.data)
        assert k == n[m]
        assert np.array_equal(la[m, :k], fine[:k])
    T = po.log_table()
    for r in range(12):
        for j in range(n[r]):
            want = L.fo_dmax(T.ctypes.data, la[r, j])
            assert dmax[r, j] == min(want, 0x7FFFFFFE)
        assert (np.diff(dmax[r, :n[r]]) <= 0).all()          # cuts shrink as alpha grows


@pytest.mark.parametrize("kind", ["small50", "genome", "dense", "c1"])
def test_cell_enumeration_matches_oracle(api, kind):
    _, flat, G = problem(kind)
    want = po.Oracle(flat).enum_cells(G)
    got = api.host_enum_cells(flat["chr_startpos"], flat["chr_bplen"], G)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("kind", ["small50", "genome", "dense"])
@pytest.mark.parametrize("rng_name", ["philox", "glibc"])
def test_block_lists_match_oracle(api, kind, rng_name):
    _, flat, _ = problem(kind)
    o = po.Oracle(flat)
    g = o.rng(rng_name, 77)
    if rng_name == "glibc":
        o.L.fo_rng_next(C.addressof(g))
    kindno = api.RNG_PHILOX if rng_name == "philox" else api.RNG_GLIBC
    n = len(flat["pos"])
    for t in range(5):
        if rng_name == "philox":
            o.L.fo_rng_set_perm(C.addressof(g), t)
        _, wb, flags = o.block_permute(g)
        got, d7 = api.host_gen_blocks(flat["chr"], flat["pos"], t, rng=kindno, seed=77)
        want = wb.copy()
        want[:, 2] = np.minimum(wb[:, 2] - wb[:, 1], n - wb[:, 0])
        assert np.array_equal(got, want)
        if rng_name == "philox":
            assert (d7 > 0) == bool(flags & 1)
        assert got[:, 2].sum() == n and (got[:, 0] == np.concatenate([[0], np.cumsum(got[:, 2])[:-1]])).all()


def _oracle_clr_matrix(o, pts, G, seed, t_list):
    """CLR of every cell on the block-permuted genomes t in t_list (restatement; stands in for the GPU)."""
    g = o.rng("philox", seed)
    rows = []
    for t in t_list:
        o.L.fo_rng_set_perm(C.addressof(g), t)
        pcls, _, _ = o.block_permute(g)
        row = []
        for s in range(len(pts)):
            start = int(pts[s]["sweep_pos"]) - int(pts[s]["sweep_pos"]) % G
            mx, _ = o.search_maxpos(int(pts[s]["chr"]), start, start + G, cls=pcls)
            row.append(mx["clr"])
        rows.append(row)
    return np.array(rows)


def test_round_replay_equals_sequential_oracle(api):
    flat, opt, _ = load_case("unfolded_n40_perm")
    G, P, seed = opt["G"], 40, 5
    o = po.Oracle(flat, mode=2)
    pts0, _ = o.scan(G)
    want, scored, _, _, wclr = o.scan_permute(pts0, P, G, o.rng("philox", seed), save_clr=True)
    full = _oracle_clr_matrix(o, pts0, G, seed, range(P + 1))
    for R in (1, 4, 7):                                   # round size must not matter
        pts = np.zeros(len(pts0), dtype=api.PT_DTYPE)
        for k in ("chr", "sweep_pos", "clr", "lalpha"):
            pts[k] = pts0[k]
        clr = np.zeros((len(pts), api.NULL_SAVE), dtype=np.float32)
        t = 0
        while t <= P:
            active = np.nonzero(pts["permute_finished"] == 0)[0]
            if len(active) == 0:
                break
            r = min(R, P + 1 - t)
            api.host_replay_round(pts, active, full[t:t + r][:, active], t, seed=seed, permute_clr=clr)
            t += r
        assert np.array_equal(pts["permute_p"], want["permute_p"])
        assert np.array_equal(pts["permute_n"], want["permute_n"])
        assert np.array_equal(pts["permute_finished"], want["finished"])
        assert np.array_equal(clr, wclr)
    assert want["finished"].sum() > 0


def test_sharded_rounds_over_gloo_two_ranks(api, tmp_path):
    """world_size 2, gloo: each rank fills the rows t % 2 == rank of the round matrix (zeros
    elsewhere), all_reduce(sum) combines them, every rank replays -> identical counts,
    equal to the sequential oracle."""
    script = os.path.join(ROOT, "tests", "gloo_worker.py")
    out = [str(tmp_path / f"r{r}.npy") for r in range(2)]
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29531", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, script, out[r]], env=dict(env, RANK=str(r)), cwd=ROOT)
             for r in range(2)]
    for p in procs:
        assert p.wait(timeout=600) == 0
    a, b = np.load(out[0]), np.load(out[1])
    assert np.array_equal(a, b)
    flat, opt, _ = load_case("unfolded_n40_perm")
    o = po.Oracle(flat, mode=2)
    pts0, _ = o.scan(opt["G"])
    want, _, _, _, _ = o.scan_permute(pts0, 30, opt["G"], o.rng("philox", 5))
    assert np.array_equal(a[0], want["permute_p"]) and np.array_equal(a[1], want["permute_n"])


def test_all_gathered_rounds_over_gloo_two_ranks(api, tmp_path):
    """The library's own sharding protocol on the CPU: rank g owns the permutations t = g (mod 2) as rows [q][n_active] of
    its chunk, ONE all_gather per round (gloo here, NCCL on the GPUs), the replay addresses row r of the round as
    (rank, local row) with the function the device kernel uses -> identical counts on both ranks, equal to the
    sequential oracle.  Rounds of 5 permutations: odd round starts and an uneven split (3 + 2 rows) every round."""
    script = os.path.join(ROOT, "tests", "gloo_worker.py")
    out = [str(tmp_path / f"g{r}.npy") for r in range(2)]
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", WORLD_SIZE="2", FSCL_TEST_ROUND="5")
    procs = [subprocess.Popen([sys.executable, script, out[r], "gather"], env=dict(env, RANK=str(r)), cwd=ROOT)
             for r in range(2)]
    for p in procs:
        assert p.wait(timeout=600) == 0
    a, b = np.load(out[0]), np.load(out[1])
    assert np.array_equal(a, b)
    flat, opt, _ = load_case("unfolded_n40_perm")
    o = po.Oracle(flat, mode=2)
    pts0, _ = o.scan(opt["G"])
    want, _, _, _, _ = o.scan_permute(pts0, 30, opt["G"], o.rng("philox", 5))
    assert np.array_equal(a[0], want["permute_p"]) and np.array_equal(a[1], want["permute_n"])


@pytest.mark.parametrize("kind", ["small50", "asc", "ascnan"])
def test_piece_maxima_bound_every_spline_value(api, kind):
    """The candidate bounds rest on pmax[l][k][c] >= spline_c(x) - null_c for every x of the knots k .. k+2^l-1
    (host-built table): checked on a dense sample of x for every class and knot; pieces with
    non-finite coefficients must bound to +inf."""
    from tests.common import problem
    tab, flat, G = problem(kind)
    coef, null_tab, step, K = flat["coef"], flat["null_tab"], flat["step"], flat["n_knots"]
    P = api.host_piece_maxima(coef, null_tab, step)
    n_cls = coef.shape[0]
    assert P.shape[1:] == (K + 2, n_cls) and P.shape[0] >= 8
    fin = np.isfinite(coef).all(axis=2) & np.isfinite(null_tab)[:, None]          # [cls][knot]
    xs = np.linspace(0.0, 1.0, 41)
    with np.errstate(all="ignore"):
        for k in range(K + 2):
            kk = min(k, K - 1)                                                       # rows K, K+1 repeat piece K-1
            x = -20.0 + (k + xs) * step
            c = coef[:, kk, :]
            c3 = c[:, 3:4] - null_tab[:, None]                                       # the device table folds the null into c3
            y = x[None, :] * (x[None, :] * (x[None, :] * c[:, 0:1] + c[:, 1:2]) + c[:, 2:3]) + c3
            ok = fin[:, kk]
            assert (P[0, k, ok][:, None] >= y[ok]).all(), (kind, k)
            assert np.isposinf(P[0, k, ~ok]).all(), (kind, k)
        # level l = running maximum over 2^l consecutive knots (clipped at the last row)
        for l in range(1, P.shape[0]):
            w = 1 << l
            for k in (0, 7, K // 2, K - 3, K + 1):
                want = P[0, k:min(k + w, K + 2), :].max(axis=0)
                assert np.array_equal(P[l, k, :], want), (kind, l, k)


def test_tile_bounds_dominate_exact_tile_sums(api):
    """The inequality behind the pruning, end to end on the CPU: for tiles wholly inside a
    candidate's cut, sum over class runs of count * (range maximum over the knots the tile spans) >= the
    exact sum of the tile's terms, for random sweep positions of a c2-shaped problem."""
    from tests.common import problem
    from oracle import pyoracle as po
    tab, flat, G = problem("mid100")
    pos, cls = flat["pos"].astype(np.int64), flat["cls"]
    coef, null_tab, step, K = flat["coef"], flat["null_tab"], flat["step"], flat["n_knots"]
    P = api.host_piece_maxima(coef, null_tab, step)[0]                 # [K+2][n_cls]
    la_tab, dmax_tab, n_tab = api.host_alpha_tables()
    la, dmax = la_tab[11][:4], dmax_tab[11][:4]                        # the four widest coarse candidates
    LT = po.log_table()

    def logt(d):
        sh = np.where(d > 0xFFFFFF, 16, np.where(d > 0xFFFF, 8, 0))
        c = np.where(d > 0xFFFFFF, 11.783502069519070, np.where(d > 0xFFFF, 5.545177444479562, 0.0))
        return np.where(sh > 0, c + LT[d >> sh], LT[d >> sh])

    rng = np.random.default_rng(4)
    T, n, checked = 1024, len(pos), 0
    for sweep in rng.integers(pos[0], pos[-1], 12):
        d = np.abs(pos - sweep)
        x0 = logt(d)
        for t in range(n // T):
            lo, hi = t * T, t * T + T
            if pos[lo] <= sweep <= pos[hi - 1]:
                continue                                                # two-sided tile: not a far tile
            dfar, dnr = int(d[lo:hi].max()), int(d[lo:hi].min())
            cnt = np.bincount(cls[lo:hi], minlength=len(null_tab))
            nz = np.nonzero(cnt)[0]
            for j in range(4):
                if dmax[j] < dfar:
                    continue                                            # the cut does not cover the tile
                x = x0[lo:hi] + la[j]
                k = np.clip(((x + 20.0) / step).astype(np.int64), 0, K - 1)
                c = coef[cls[lo:hi], k]
                exact = float(np.sum(x * (x * (x * c[:, 0] + c[:, 1]) + c[:, 2]) + (c[:, 3] - null_tab[cls[lo:hi]])))
                ka = int(np.clip((logt(np.array([dnr]))[0] + la[j] + 20.0) / step, 0, K + 1))
                kb = int(np.clip((logt(np.array([dfar]))[0] + la[j] + 20.0) / step, 0, K + 1))
                ub = float(np.sum(cnt[nz] * P[max(ka - 1, 0):kb + 2][:, nz].max(axis=0)))   # +-1 knot: device knot map vs this one
                assert ub >= exact - 1e-9 * (1.0 + abs(exact)), (int(sweep), t, j, ub, exact)
                checked += 1
    assert checked > 300


def test_grouped_block_swaps_equal_the_sequential_swap_loop():
    """permute_kernel applies consecutive blocks that touch no common site as one set of swaps and a self-overlapping
    block through a closed form (tools/proto/block_groups_proto.py states its rule): same array as snp_block_permute's
    element-by-element loop (scan-chromosome.c:365-384) on random lists where most blocks overlap something."""
    sys.path.insert(0, os.path.join(ROOT, "tools", "proto"))
    import block_groups_proto as bg
    rng = np.random.default_rng(7)
    sizes = []
    for trial in range(200):
        n = int(rng.integers(50, 3000))
        blocks = []
        for _ in range(int(rng.integers(1, 120))):
            L = int(rng.integers(0, max(2, n // int(rng.integers(2, 40)))))
            i, j = int(rng.integers(0, n - L + 1)), int(rng.integers(0, n - L + 1))
            if rng.random() < 0.05:
                j = i
            blocks.append((i, j, L))
        a = rng.integers(0, 1000, n)
        assert np.array_equal(bg.apply_sequential(a, blocks), bg.apply_grouped(a, blocks)), trial
        sizes += bg.group_sizes(blocks)
    assert min(sizes) == 1 and max(sizes) > 4               # both single blocks and real groups occurred
