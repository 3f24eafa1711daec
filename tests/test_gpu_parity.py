"""GPU parity tests: CUDA path (through the C-ABI) vs the CPU oracle."""
import numpy as np
import pytest

from tests.common import clr_close, problem

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api(built):
    from fscl_b200 import api as a
    if a.load_library().fscl_device_count() <= 0:
        pytest.fail("no CUDA device visible: the gpu tests must not pass without one")
    return a


def _oracle(flat, mode=2):
    from oracle import pyoracle as po
    return po.Oracle(flat, mode=mode)


@pytest.mark.parametrize("kind", ["small50", "c1", "asc", "asc1", "ascnan", "genome", "dense", "invariant", "folded"])
def test_grid_eval_matches_oracle(api, kind):
    tab, flat, G = problem(kind)
    sc = api.Scanner(flat)
    o = _oracle(flat)
    rng = np.random.default_rng(1)
    n_chr = len(flat["chr_start"])
    chrs = rng.integers(0, n_chr, 200).astype(np.int32)
    lo, hi = flat["chr_startpos"][chrs], flat["chr_bplen"][chrs]
    pos = (lo + rng.random(200) * (hi - lo + 2000) - 1000).astype(np.int32)
    pos[:50] = flat["pos"][flat["chr_start"][chrs[:50]] + rng.integers(0, flat["chr_nsnps"][chrs[:50]])]  # on a SNP
    got = sc.grid_eval(chrs, pos)
    for i in range(len(pos)):
        w = o.grid_eval(int(chrs[i]), int(pos[i]))
        for a, b in (("nearest_snp", "nearest"), ("sweep_pos", "sweep_pos"), ("window_start", "wstart"),
                     ("window_end", "wend"), ("n_snps", "n_snps")):
            assert got[i][a] == w[b], (kind, i, a)
        assert clr_close(got[i]["clr"], w["clr"]), (kind, i, got[i]["clr"], w["clr"])
        assert got[i]["lalpha"] == w["lalpha"], (kind, i)
        assert abs(got[i]["null_logl"] - w["null_logl"]) <= 1e-12 * abs(w["null_logl"])


@pytest.mark.parametrize("kind", ["small50", "c1", "asc", "asc1", "ascnan", "genome", "dense", "invariant", "folded"])
def test_scan_chromosome_matches_oracle(api, kind):
    tab, flat, G = problem(kind)
    sc = api.Scanner(flat)
    got = sc.scan_chromosome(G)
    o = _oracle(flat)
    want, _ = o.scan(G)
    assert len(got) == len(want)
    assert np.array_equal(got["chr"], want["chr"])
    assert np.array_equal(got["sweep_pos"], want["sweep_pos"])
    assert np.array_equal(got["window_start"], want["wstart"])
    assert np.array_equal(got["window_end"], want["wend"])
    assert clr_close(got["clr"], want["clr"]).all()
    assert np.array_equal(got["lalpha"], want["lalpha"])
    # output records are byte-identical
    assert sc.format_records(got) == o.format_records(want)


@pytest.mark.parametrize("kind", ["small50", "genome", "dense"])
@pytest.mark.parametrize("rng_name", ["philox", "glibc"])
def test_block_permutation_bit_exact(api, kind, rng_name):
    tab, flat, G = problem(kind)
    sc = api.Scanner(flat)
    o = _oracle(flat)
    kindno = api.RNG_PHILOX if rng_name == "philox" else api.RNG_GLIBC
    g = o.rng(rng_name, 1234)
    if rng_name == "glibc":
        o.L.fo_rng_next(__import__("ctypes").addressof(g))      # usleep burn
    for t in range(4):
        if rng_name == "philox":
            o.L.fo_rng_set_perm(__import__("ctypes").addressof(g), t)
        want, wblocks, flags = o.block_permute(g)
        got, gblocks = sc.block_permute(t, rng=kindno, seed=1234)
        wb = wblocks.copy()
        wb[:, 2] = np.minimum(wb[:, 2] - wb[:, 1], len(want) - wb[:, 0])      # (i, j, k) -> (i, j, len)
        assert np.array_equal(gblocks, wb), (kind, rng_name, t)
        assert np.array_equal(got, want), (kind, rng_name, t)
        assert not np.array_equal(got, flat["cls"])


@pytest.mark.parametrize("kind,n_perm", [("small50", 60), ("dense", 40), ("asc", 30), ("ascnan", 25)])
@pytest.mark.parametrize("rng_name", ["philox", "glibc"])
def test_scan_permute_counts_bit_exact(api, kind, n_perm, rng_name):
    tab, flat, G = problem(kind)
    sc = api.Scanner(flat)
    o = _oracle(flat)
    obs = sc.scan_chromosome(G)
    want0, _ = o.scan(G)
    g = o.rng(rng_name, 99)
    want, scored, flags, ne, wclr = o.scan_permute(want0, n_perm, G, g, save_clr=True)
    kindno = api.RNG_PHILOX if rng_name == "philox" else api.RNG_GLIBC
    got, st, gclr = sc.scan_permute(obs, n_perm, G, rng=kindno, seed=99, save_clr=True)
    assert np.array_equal(got["permute_p"], want["permute_p"])
    assert np.array_equal(got["permute_n"], want["permute_n"])
    assert np.array_equal(got["permute_finished"], want["finished"])
    assert st["n_scored"] == scored
    assert want["finished"].sum() > 0, "case must exercise pruning"
    assert np.allclose(gclr, wclr, rtol=1e-6, atol=1e-6)
    assert sc.format_records(got, n_perm) == o.format_records(want, n_perm)


def test_round_size_invariance(api):
    tab, flat, G = problem("small50")
    sc = api.Scanner(flat)
    obs = sc.scan_chromosome(G)
    a, _, _ = sc.scan_permute(obs, 50, G, rng=api.RNG_PHILOX, seed=5, round_perms=1)
    b, _, _ = sc.scan_permute(obs, 50, G, rng=api.RNG_PHILOX, seed=5, round_perms=7)
    c, _, _ = sc.scan_permute(obs, 50, G, rng=api.RNG_PHILOX, seed=5)
    for x in (b, c):
        assert np.array_equal(a["permute_p"], x["permute_p"]) and np.array_equal(a["permute_n"], x["permute_n"])


@pytest.mark.parametrize("kind", ["small50", "genome", "dense", "mid100"])
def test_device_block_draws_equal_host_and_oracle(api, kind):
    """gen_blocks_kernel (Philox on the device) vs the host generator vs the oracle, bit for bit."""
    import ctypes as C
    tab, flat, G = problem(kind)
    sc = api.Scanner(flat)
    o = _oracle(flat)
    g = o.rng("philox", 4242)
    n = len(flat["pos"])
    for t in (0, 1, 2, 17, 1000):
        dev, flags = sc.device_gen_blocks(t, seed=4242)
        host, d7 = api.host_gen_blocks(flat["chr"], flat["pos"], t, seed=4242)
        assert not (flags & 6), "no ambiguous truncation / overflow expected"
        assert np.array_equal(dev, host)
        assert bool(flags & 1) == (d7 > 0)
        o.L.fo_rng_set_perm(C.addressof(g), t)
        _, wb, _ = o.block_permute(g)
        want = wb.copy()
        want[:, 2] = np.minimum(wb[:, 2] - wb[:, 1], n - wb[:, 0])
        assert np.array_equal(dev, want)


def test_device_rounds_equal_host_rounds(api, monkeypatch):
    """Counting / pruning / compaction on the device vs the host replay of the same rounds."""
    tab, flat, G = problem("small50")
    sc = api.Scanner(flat)
    obs = sc.scan_chromosome(G)
    a, sta, ca = sc.scan_permute(obs, 70, G, rng=api.RNG_PHILOX, seed=11, save_clr=True)
    monkeypatch.setenv("FSCL_B200_HOST_ROUNDS", "1")
    sc.reset_stats()
    b, stb, cb = sc.scan_permute(obs, 70, G, rng=api.RNG_PHILOX, seed=11, save_clr=True)
    for k in ("permute_p", "permute_n", "permute_finished"):
        assert np.array_equal(a[k], b[k])
    assert np.array_equal(ca, cb)
    assert sta["n_scored"] == stb["n_scored"] and a["permute_finished"].sum() > 0


def test_permutation_test_resumes_bit_exactly(api):
    """Checkpoint / resume: stopping after a few rounds and continuing from the
    snapshot gives the counts, flags and stored CLRs of an uninterrupted run."""
    tab, flat, G = problem("small50")
    sc = api.Scanner(flat)
    obs = sc.scan_chromosome(G)
    full, _, clr_full = sc.scan_permute(obs, 60, G, save_clr=True, rng=api.RNG_PHILOX, seed=12, round_perms=8)
    p1, _, c1, t1 = sc.scan_permute_rounds(obs, 60, G, save_clr=True, max_rounds=2, rng=api.RNG_PHILOX, seed=12,
                                           round_perms=8)
    assert t1 == 16 and int(p1["permute_n"].max()) == 16
    p2, _, c2, t2 = sc.scan_permute_rounds(p1, 60, G, clr=c1, start_t=t1, max_rounds=3, rng=api.RNG_PHILOX, seed=12,
                                           round_perms=8)
    p3, _, c3, t3 = sc.scan_permute_rounds(p2, 60, G, clr=c2, start_t=t2, rng=api.RNG_PHILOX, seed=12, round_perms=8)
    assert t2 == 40 and t3 == 61
    for k in ("permute_p", "permute_n", "permute_finished"):
        assert np.array_equal(p3[k], full[k]), k
    assert np.array_equal(c3, clr_full)


def test_create_rejects_bad_input_and_releases_the_context(api):
    """Errors are return codes with a message; a context that fails half-way frees what it owns."""
    tab, flat, G = problem("small50")
    bad = dict(flat)
    bad["cls"] = flat["cls"].copy()
    bad["cls"][17] = len(flat["null_tab"]) + 3
    for _ in range(3):
        with pytest.raises(api.FsclError, match="class index out of range"):
            api.Scanner(bad)
    sc = api.Scanner(flat)                      # the library is still usable afterwards
    assert len(sc.scan_chromosome(G)) > 0
