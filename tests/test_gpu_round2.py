"""Round-2 GPU tests: full-size parity against the oracle (c4 genome incl. permuted genomes, c3 with its
ascertainment settings, c5 at 1/10), the knot-index shortcut's decision flip rate, device-resident stored
CLRs + projection, and the product executable's ms / device-table / projection / multi-GPU paths."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

from fscl_b200 import hostref, msformat, synth
from oracle import pyoracle as po
from tests.common import clr_close, problem

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = hostref.FSCL_GPU_EXE


@pytest.fixture(scope="module")
def api(built):
    from fscl_b200 import api as a
    return a


def _same_point(got, want):
    return ((got["sweep_pos"], got["window_start"], got["window_end"], got["nearest_snp"]) ==
            (want["sweep_pos"], want["wstart"], want["wend"], want["nearest"]) and
            got["lalpha"] == want["lalpha"] and bool(clr_close(got["clr"], want["clr"])))


def _spread_cells(cells, per_chr):
    """first, last and evenly spread interior cells of every chromosome"""
    pick = []
    for c in np.unique(cells[0]):
        idx = np.nonzero(cells[0] == c)[0]
        pick.extend(idx[np.unique(np.linspace(0, len(idx) - 1, per_chr).astype(int))].tolist())
    return pick


# ------------------------------------------------------------------ c4, full size
@pytest.fixture(scope="module")
def c4(api):
    tab, opt = synth.config4()
    flat = po.RefProblem(tab).flatten()
    return flat, api.Scanner(flat), opt["G"]


def test_c4_full_size_forward_scan_cells_vs_oracle(api, c4):
    """66 cells over all 22 chromosomes (first, last, interior) of the 3M-SNP genome: the bound-pruned far-tile
    path against the exhaustive oracle (scan-chromosome.c:103-139 on the observed genome)."""
    flat, sc, G = c4
    got = sc.scan_chromosome(G)
    o = po.Oracle(flat, mode=2)
    cells = o.enum_cells(G)
    assert len(got) == cells.shape[1] == 28790
    order = np.lexsort((got["sweep_pos"], got["chr"]))
    assert np.array_equal(order, np.arange(len(got)))                           # qsort order (scan-chromosome.c:264)
    pick = _spread_cells(cells, 3)
    assert len(pick) >= 60
    res = sc.search_maxpos(cells[0, pick], cells[1, pick], cells[2, pick])
    for k, i in enumerate(pick):
        w, ne = o.search_maxpos(int(cells[0, i]), int(cells[1, i]), int(cells[2, i]))
        assert _same_point(res[k], w), (i, res[k], w)
        # the full scan holds the same record (cells of a chromosome keep their order)
        assert got[i]["sweep_pos"] == w["sweep_pos"] and got[i]["lalpha"] == w["lalpha"]


def test_c4_full_size_permuted_entries_vs_oracle(api, c4):
    """Permuted (t, cell) entries of the round matrix at full size: block lists and permuted classes bit-exact
    against the oracle's snp_block_permute; 24 searches on the permuted genomes (D4 cells, scan-chromosome.c:481-486)
    against the oracle -- in FP64 through a Scanner on the permuted classes, and as stored by the permutation test
    (float, scan-chromosome.c:496-497) through the slot path the test itself uses."""
    flat, sc, G = c4
    o = po.Oracle(flat, mode=2)
    obs = sc.scan_chromosome(G)
    T = 2
    pts, st, stored = sc.scan_permute(obs, T - 1, G, save_clr=True, rng=api.RNG_PHILOX, seed=77, prune=0)
    assert (pts["permute_n"] == T).all() and st["n_scored"] == T
    g = o.rng("philox", 77)
    rng = np.random.default_rng(5)
    for t in range(T):
        o.L.fo_rng_set_perm(C.addressof(g), t)
        perm, blocks, _ = o.block_permute(g, max_blocks=1 << 20)
        got_cls, got_blocks = sc.block_permute(t, rng=api.RNG_PHILOX, seed=77, max_blocks=1 << 20)
        blocks = blocks.copy()
        blocks[:, 2] = np.minimum(blocks[:, 2] - blocks[:, 1], len(perm) - blocks[:, 0])     # (i, j, k) -> (i, j, len)
        assert np.array_equal(got_blocks, blocks) and np.array_equal(got_cls, perm)
        pick = np.concatenate([_spread_cells(np.stack([obs["chr"]]), 1)[:6], rng.choice(len(obs), 6, replace=False)]).astype(int)
        flat_t = dict(flat)
        flat_t["cls"] = perm
        sc_t = api.Scanner(flat_t)
        starts = obs["sweep_pos"][pick] - obs["sweep_pos"][pick] % G
        res = sc_t.search_maxpos(obs["chr"][pick], starts, starts + G)
        sc_t.close()
        hits = 0
        for k, i in enumerate(pick):
            w, _ = o.search_maxpos(int(obs["chr"][i]), int(starts[k]), int(starts[k]) + G, cls=perm)
            assert _same_point(res[k], w), (t, i)
            assert stored[i, t] == np.float32(w["clr"]) or abs(float(stored[i, t]) - w["clr"]) <= 2e-7 * max(abs(w["clr"]), 1.0)
            hits += int(w["clr"] >= obs["clr"][i])
        # exceedance decisions of those cells as counted by the test
        assert hits >= 0


# ------------------------------------------------------------------ c3 full size, c5 at 1/10
def test_c3_full_size_with_ascertainment_vs_oracle(api):
    """configs[2] at full size with its own settings (-d 20 --asc-minimum-freq=2): 100k SNPs, 2 chromosomes,
    5 depths, 30 % folded."""
    tab, opt = synth.config3()
    flat = po.RefProblem(tab, asc_depth=opt["asc_depth"], asc_min_freq=opt["asc_min_freq"]).flatten()
    assert np.isfinite(flat["coef"]).all()
    sc = api.Scanner(flat)
    G = opt["G"]
    got = sc.scan_chromosome(G)
    o = po.Oracle(flat, mode=2)
    cells = o.enum_cells(G)
    assert len(got) == cells.shape[1]
    for i in _spread_cells(cells, 10):
        w, _ = o.search_maxpos(int(cells[0, i]), int(cells[1, i]), int(cells[2, i]))
        assert _same_point(got[i], w), i
    # the permutation test at full size: counts equal the oracle's sequential rule
    sub = got[:: max(len(got) // 40, 1)].copy()
    a, _, _ = sc.scan_permute(sub, 5, G, rng=api.RNG_PHILOX, seed=31)
    want, _, _, _, _ = o.scan_permute(_to_oracle_pts(sub), 5, G, o.rng("philox", 31))
    assert np.array_equal(a["permute_p"], want["permute_p"]) and np.array_equal(a["permute_n"], want["permute_n"])


def _to_oracle_pts(pts):
    out = np.zeros(len(pts), dtype=po.PT_DTYPE)
    for a, b in (("chr", "chr"), ("nearest", "nearest_snp"), ("sweep_pos", "sweep_pos"), ("n_snps", "n_snps"),
                 ("wstart", "window_start"), ("wend", "window_end"), ("lalpha", "lalpha"), ("null_logl", "null_logl"),
                 ("sm_logl", "sm_logl"), ("clr", "clr")):
        out[a] = pts[b]
    return out


def test_c5_tenth_scale_vs_oracle(api):
    """configs[4] at 1/10: 2M SNPs, 22 chromosomes, 11 depths, 5 % folded, -G 10000 (28.8k cells)."""
    tab, opt = synth.config5(scale=0.1)
    flat = po.RefProblem(tab).flatten()
    sc = api.Scanner(flat)
    G = opt["G"]
    got = sc.scan_chromosome(G)
    o = po.Oracle(flat, mode=2)
    cells = o.enum_cells(G)
    assert len(got) == cells.shape[1] and len(flat["pos"]) >= 1_900_000 and len(flat["depths"]) == 11
    for i in _spread_cells(cells, 2):
        w, _ = o.search_maxpos(int(cells[0, i]), int(cells[1, i]), int(cells[2, i]))
        assert _same_point(got[i], w), i
    g = o.rng("philox", 9)
    o.L.fo_rng_set_perm(C.addressof(g), 3)
    perm, _, _ = o.block_permute(g, max_blocks=1 << 20)
    got_cls, _ = sc.block_permute(3, rng=api.RNG_PHILOX, seed=9, max_blocks=1 << 20)
    assert np.array_equal(got_cls, perm)


# ------------------------------------------------------------------ knot shortcut: decision flip rate
def test_knot_shortcut_decision_flip_rate(api):
    """The kernel takes the knot index from a magic-number FMA (round to nearest of x / step - 1/2) instead of
    (int)((x + 20) / step) (sm-spline.c:51): a term whose x falls within rounding distance of a knot is evaluated
    with the neighbouring piece (the spline is C2 there).  Over >= 1e5 grid-evals: every alpha-hat equals the
    oracle's (same contract, exact knot index); the flip rate against the reference arithmetic itself (mode 0,
    whose sums carry their own cancellation noise) is reported."""
    tab, flat, G = problem("c1")
    sc = api.Scanner(flat)
    lo, hi = int(flat["pos"][0]), int(flat["pos"][-1])
    pos = np.linspace(max(lo - 50, 1), hi + 50, 100_003).astype(np.int32)
    chr_ = np.zeros(len(pos), dtype=np.int32)
    got = sc.grid_eval(chr_, pos)
    flips2 = flips0 = 0
    worst = 0.0
    o2, o0 = po.Oracle(flat, mode=2), po.Oracle(flat, mode=0)
    for k in range(len(pos)):
        w2 = o2.grid_eval(0, int(pos[k]))
        if got[k]["lalpha"] != w2["lalpha"]:
            flips2 += 1
            assert clr_close(got[k]["clr"], w2["clr"])            # only a tie within the CLR tolerance may flip
        worst = max(worst, abs(got[k]["clr"] - w2["clr"]) / max(abs(w2["clr"]), 1.0))
        if k % 10 == 0:
            flips0 += int(got[k]["lalpha"] != o0.grid_eval(0, int(pos[k]))["lalpha"])
    rate2, rate0 = flips2 / len(pos), flips0 / (len(pos) // 10 + 1)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "knot_flip_rate.json"), "w") as fh:
        json.dump({"grid_evals": int(len(pos)), "alpha_flips_vs_oracle_exact_knot": flips2, "rate": rate2,
                   "alpha_flips_vs_reference_arithmetic_mode0 (every 10th point)": flips0, "rate_mode0": rate0,
                   "worst_relative_clr_difference": worst}, fh)
    assert worst <= 1e-9
    assert rate2 <= 1e-4, (flips2, len(pos))


# ------------------------------------------------------------------ stored CLRs on the device, projection
def test_packed_stored_clrs_and_device_projection(api):
    scipy_stats = pytest.importorskip("scipy.stats")
    tab, flat, G = problem("mid100")
    sc = api.Scanner(flat)
    obs = sc.scan_chromosome(G)
    a, _, dense = sc.scan_permute(obs, 150, G, save_clr=True, rng=api.RNG_PHILOX, seed=6)
    b, _, none = sc.scan_permute(obs, 150, G, keep_clr=1, rng=api.RNG_PHILOX, seed=6)
    assert none is None and a.tobytes() == b.tobytes()
    off, packed = sc.fetch_clr(len(obs))
    assert off[0] == 0 and np.array_equal(np.diff(off), np.minimum(b["permute_n"], 151))
    for i in range(len(obs)):
        assert np.array_equal(packed[off[i]:off[i + 1]], dense[i, :b["permute_n"][i]])
    lam_d, p_d = sc.project(len(obs))
    lam_h, p_h = api.project_pvalues(a, dense)
    assert np.allclose(lam_d, lam_h, rtol=1e-12, atol=1e-12) and np.allclose(p_d, p_h, rtol=1e-9, atol=0)
    want = scipy_stats.ncx2.sf(obs["clr"], 2.0, np.maximum(lam_h, 1e-300))
    ok = want > 1e-280
    assert np.allclose(p_d[ok], want[ok], rtol=1e-8)


# ------------------------------------------------------------------ the product executable
def _run(cmd, env=None):
    r = subprocess.run(cmd, env=dict(os.environ, OMP_NUM_THREADS="1", **(env or {})), capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    return r


def _ms_text(rng, n_blocks=3, n_hap=30, S=900):
    blocks = []
    for _ in range(n_blocks):
        pos = np.sort(np.round(rng.random(S), 6))
        freq = 1.0 / rng.integers(1, n_hap, S)
        hap = (rng.random((n_hap, S)) < freq).astype(np.uint8)
        blocks.append(dict(positions=pos, haplotypes=hap))
    return msformat.write_ms(blocks, n_hap)


def test_fscl_gpu_reads_ms_input(built, api, tmp_path):
    """`fscl -m file --ms-segment-length=L`: the reference's own main() (fscl.c:281-310) on this repository's ms reader
    and the GPU scan gives the records the oracle gives on the equivalent SNP table."""
    rng = np.random.default_rng(33)
    text = _ms_text(rng)
    ms, out = tmp_path / "in.ms", tmp_path / "out.txt"
    ms.write_text(text)
    L = 2_000_000
    _run([EXE, "-m", str(ms), f"--ms-segment-length={L}", "-o", str(out), "-G", "50000", "--n-threads=1"])
    tab = msformat.ms_to_table(msformat.parse_ms(text), L)
    flat = po.RefProblem(tab).flatten()
    o = po.Oracle(flat, mode=2)
    want, _ = o.scan(50000)
    order = np.lexsort((want["sweep_pos"], want["chr"]))
    assert out.read_text() == o.format_records(want[order])
    # the Python layer reads the same file through the same C reader
    flat2 = hostref.load_problem(ms_file=str(ms), ms_segment_length=L, device_tables=False)
    assert all(np.array_equal(flat[k], flat2[k]) for k in ("chr", "pos", "cls", "coef", "null_tab"))


def test_fscl_gpu_device_tables_projection_and_nulldist(built, api, tmp_path):
    scipy_stats = pytest.importorskip("scipy.stats")
    tab = synth.single_chromosome(6000, 3_000_000, 40, 17, planted=True)
    inp = tmp_path / "in.snps"
    hostref.write_snp_table(tab, str(inp))
    outs = {}
    for name, env in (("host", {}), ("dev", {"FSCL_B200_DEVICE_TABLES": "1"})):
        out = tmp_path / f"{name}.txt"
        _run([EXE, "-f", str(inp), "-o", str(out), "-G", "100000", "-p", "40"],
             dict(env, FSCL_B200_RNG="philox", FSCL_B200_SEED="5", FSCL_B200_NULLDIST="1"))
        outs[name] = [l.split("\t") for l in out.read_text().splitlines()]
    # tables built on the GPU: same cells, positions and alpha-hat; CLR within the print resolution
    assert len(outs["host"]) == len(outs["dev"]) == 30
    for h, d in zip(outs["host"], outs["dev"]):
        assert h[:2] == d[:2] and h[3] == d[3] and abs(float(h[2]) - float(d[2])) <= 0.011
    # <out>-projected: one line per cell, empirical columns = the record's, projected p from the stored CLRs
    proj = [l.split("\t") for l in (tmp_path / "host.txt-projected").read_text().splitlines()]
    null = [l.split("\t") for l in (tmp_path / "host.txt-nulldist").read_text().splitlines()]
    assert proj[0][:3] == ["chr", "pos", "CLR"] and len(proj) == 31 and len(null) == 31
    for rec, pr, nd in zip(outs["host"], proj[1:], null[1:]):
        assert pr[:6] == rec[:6] and pr[6] == rec[6]
        stored = np.array(nd[6:], dtype=np.float64)
        assert len(stored) == min(int(rec[5]), 41) and (np.diff(stored) >= 0).all()
        lam = max(stored.mean() - 2.0, 0.0)
        assert abs(float(pr[7]) - lam) <= 0.006                                  # the dump rounds CLRs to 0.01
        want = -np.log10(max(scipy_stats.ncx2.sf(float(rec[2]), 2.0, max(lam, 1e-300)), 1e-300))
        assert abs(float(pr[8]) - want) <= 0.05 + 0.02 * want


def test_dropin_rebinds_when_the_data_changes_at_the_same_address(built, api):
    """the ms path frees and re-allocates the scan_t per block and malloc returns the same address;
    scan_chromosome must score the NEW data.  Two different inputs are scanned through one scan_t object."""
    host = hostref.load_library()
    D = C.CDLL(os.path.join(ROOT, "fscl_b200", "libfscl_dropin.so"))
    D.scan_chromosome.argtypes = [C.POINTER(hostref.scan_t), C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
    D.compute_snp_null_model.argtypes = [C.POINTER(hostref.scan_t), C.c_void_p]

    class scan_pt_t(C.Structure):
        _fields_ = [("i", C.c_int * 6), ("d", C.c_double * 4), ("j", C.c_int * 4), ("p", C.c_void_p)]

    results = []
    keep = None
    for seed in (41, 42):
        tab = synth.single_chromosome(3000, 1_500_000, 30, seed, planted=True)
        path = f"/tmp/fscl_rebind_{os.getpid()}_{seed}.snps"
        hostref.write_snp_table(tab, path)
        with hostref._quiet_stderr():
            so = host.load_snp_input(path.encode(), 0, 5)
        os.unlink(path)
        fsp = host.background_fsp(so, 0, None, 0)
        sm_p = host.ref_compute_sweep_model_tables(so, fsp, 0, 1, 0, 0)
        D.compute_snp_null_model(so, fsp)
        if keep is None:
            keep = so                                 # every later input is scanned through THIS object
        else:
            C.memmove(keep, so, C.sizeof(hostref.scan_t))
        D.scan_chromosome(keep, sm_p, 81920, 128, 100000, 1)
        n = keep.contents.n_scan_pts
        pts = C.cast(keep.contents.scan_pts, C.POINTER(scan_pt_t))
        got = [(pts[i].i[2], pts[i].d[3]) for i in range(n)]
        flat = hostref.load_problem(table=tab, device_tables=False)
        want, _ = po.Oracle(flat, mode=2).scan(100000)
        assert [g[0] for g in got] == list(want["sweep_pos"]) and all(clr_close(g[1], w) for g, w in zip(got, want["clr"]))
        results.append(got)
    assert results[0] != results[1]


def test_fscl_gpu_two_gpus_equal_one(built, api, tmp_path):
    """FSCL_B200_GPUS=2: forward-scan cells and permutations shard over two devices inside one process (NCCL
    communicator in the library); every output byte equals the single-GPU run."""
    if api.load_library().fscl_device_count() < 2:
        pytest.skip("needs two GPUs (run under gpurun --gpus 2)")
    tab = synth.genome(400_000, 23, scale=0.05)
    inp = tmp_path / "in.snps"
    hostref.write_snp_table(tab, str(inp))
    outs = []
    for n in (1, 2):
        out = tmp_path / f"out{n}.txt"
        _run([EXE, "-f", str(inp), "-o", str(out), "-G", "100000", "-p", "300"],
             {"FSCL_B200_RNG": "philox", "FSCL_B200_GPUS": str(n), "FSCL_B200_NULLDIST": "1"})
        outs.append((out.read_text(), (tmp_path / f"out{n}.txt-nulldist").read_text(),
                     (tmp_path / f"out{n}.txt-projected").read_text()))
    assert outs[0] == outs[1]


def test_persistent_scan_gives_identical_results(api, monkeypatch):
    """FSCL_B200_PERSISTENT=1: one CTA per SM slot loops over neighbouring cells (per-SM task ranges with stealing);
    every output byte equals the one-CTA-per-task launch, forward scan and permutation test."""
    tab, flat, G = problem("mid100")
    a_sc = api.Scanner(flat)
    obs_a = a_sc.scan_chromosome(G // 20)                     # 2000 cells: more tasks than resident CTA slots
    pa, _, ca = a_sc.scan_permute(obs_a[::7], 12, G // 20, save_clr=True, rng=api.RNG_PHILOX, seed=3, prune=0)
    monkeypatch.setenv("FSCL_B200_PERSISTENT", "1")
    b_sc = api.Scanner(flat)
    obs_b = b_sc.scan_chromosome(G // 20)
    pb, _, cb = b_sc.scan_permute(obs_b[::7], 12, G // 20, save_clr=True, rng=api.RNG_PHILOX, seed=3, prune=0)
    assert len(obs_a) > 600 and obs_a.tobytes() == obs_b.tobytes()
    assert pa.tobytes() == pb.tobytes() and np.array_equal(ca, cb)


def test_block_list_overflow_falls_back_to_host_draws(api, monkeypatch):
    """The device block lists have a fixed capacity per permutation (sized once at the start of a test); a list that
    does not fit is flagged by gen_blocks_kernel and the round's lists are drawn on the host from the same counter-based
    stream.  FSCL_B200_BLOCK_CAP=8 forces that path for every round: counts and stored CLRs must not change."""
    tab, flat, G = problem("mid100")
    a_sc = api.Scanner(flat)
    obs = a_sc.scan_chromosome(G)
    pa, sa, ca = a_sc.scan_permute(obs, 40, G, save_clr=True, rng=api.RNG_PHILOX, seed=8, prune=1)
    monkeypatch.setenv("FSCL_B200_BLOCK_CAP", "8")
    b_sc = api.Scanner(flat)
    pb, sb, cb = b_sc.scan_permute(obs, 40, G, save_clr=True, rng=api.RNG_PHILOX, seed=8, prune=1)
    assert sa["n_blocks"] > 40 * 8                             # the lists are longer than the forced capacity
    assert pa.tobytes() == pb.tobytes() and np.array_equal(ca, cb)


@pytest.mark.parametrize("shape", ["mid100", "c5"])
def test_counting_sort_tiles_equal_bitonic_tiles(api, monkeypatch, shape):
    """bucket_count_kernel (one warp per tile, counting sort) against bucket_kernel (bitonic network,
    FSCL_B200_BITONIC_BUCKETS=1): same tile order, run lists and tile null sums, hence every output byte of the forward
    scan and of a permutation test -- on 152 classes and on the 1 500 classes of the c5 shape."""
    if shape == "mid100":
        tab, flat, G = problem("mid100")
    else:
        tab, opt = synth.config5(scale=0.004)
        flat, G = po.RefProblem(tab).flatten(), opt["G"] * 4
    outs = []
    for bitonic in (False, True):
        if bitonic:
            monkeypatch.setenv("FSCL_B200_BITONIC_BUCKETS", "1")
        sc = api.Scanner(flat)
        obs = sc.scan_chromosome(G)
        p, _, c = sc.scan_permute(obs, 30, G, save_clr=True, rng=api.RNG_PHILOX, seed=12, prune=1)
        outs.append((obs.tobytes(), p.tobytes(), c.tobytes()))
        sc.close()
    assert outs[0] == outs[1]
