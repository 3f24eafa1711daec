"""Python host layer over the C-ABI of ``libfscl_b200.so`` (include/fscl_b200.h).

Mirrors the reference's scan entry points (slowkoni/fscl ``fscl.h:103-116``):

* ``Scanner.scan_chromosome(G)``      <- ``scan_chromosome``  scan-chromosome.c:228
* ``Scanner.scan_permute(pts, ...)``  <- ``scan_permute``     scan-chromosome.c:582
* ``Scanner.grid_eval(chr, pos)``     <- ``init_scan_result`` + ``search_maxalpha``
* ``format_records(...)``             <- ``scan_output``      scan-chromosome.c:666

There is no CPU path: construction fails when the CUDA library is missing or no
sm_100 device is present.
"""
from __future__ import annotations

import ctypes as C
import math
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FSCL_B200_LIB", os.path.join(HERE, "libfscl_b200.so"))
NULL_SAVE = 10000

RNG_GLIBC, RNG_PHILOX, RNG_LIBC = 0, 1, 2
DEFAULT_SEED = 0xFD821A6          # srand() at fscl.c:135


class FsclError(RuntimeError):
    pass


class fscl_problem_t(C.Structure):
    _fields_ = [("n_snps", C.c_int32), ("chr", C.c_void_p), ("pos", C.c_void_p), ("cls", C.c_void_p),
                ("n_chr", C.c_int32), ("chr_start", C.c_void_p), ("chr_nsnps", C.c_void_p),
                ("chr_startpos", C.c_void_p), ("chr_bplen", C.c_void_p), ("n_cls", C.c_int32),
                ("n_knots", C.c_int32), ("step", C.c_double), ("coef", C.c_void_p), ("null_tab", C.c_void_p),
                ("log_table", C.c_void_p), ("eval_range", C.c_int32), ("bp_resl", C.c_int32),
                ("device", C.c_int32), ("max_round_perms", C.c_int32)]


class fscl_permute_opts_t(C.Structure):
    _fields_ = [("n_permute", C.c_int32), ("permute_nbp", C.c_double), ("large_grid_sp", C.c_int32),
                ("scan_width_mb", C.c_double), ("rng_kind", C.c_int32), ("seed", C.c_uint64),
                ("prune", C.c_int32), ("first_perm", C.c_int32), ("perm_stride", C.c_int32),
                ("round_perms", C.c_int32), ("save_clr", C.c_int32)]


class fscl_stats_t(C.Structure):
    _fields_ = [("n_scored", C.c_int64), ("n_tasks", C.c_int64), ("n_grid_evals", C.c_int64),
                ("n_rounds", C.c_int64), ("n_blocks", C.c_int64), ("d7_hits", C.c_int32),
                ("bound_failures", C.c_int32), ("score_kernel_ms", C.c_double), ("permute_kernel_ms", C.c_double),
                ("score_launches", C.c_int64), ("permute_launches", C.c_int64),
                ("kernel_launches", C.c_int64), ("near_ties_alpha", C.c_int64), ("near_ties_position", C.c_int64),
                ("fine_bound_failures", C.c_int64)]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


PT_DTYPE = np.dtype([("chr", "<i4"), ("nearest_snp", "<i4"), ("sweep_pos", "<i4"), ("n_snps", "<i4"),
                     ("window_start", "<i4"), ("window_end", "<i4"), ("lalpha", "<f8"), ("null_logl", "<f8"),
                     ("sm_logl", "<f8"), ("clr", "<f8"), ("permute_n", "<i4"), ("permute_p", "<i4"),
                     ("permute_finished", "<i4"), ("reserved", "<i4")])
assert PT_DTYPE.itemsize == 72

EXPORTS = ["fscl_last_error", "fscl_device_count", "fscl_create", "fscl_destroy", "fscl_enum_cells",
           "fscl_grid_eval", "fscl_search_maxpos", "fscl_scan_chromosome", "fscl_block_permute",
           "fscl_scan_permute", "fscl_perm_begin", "fscl_perm_round_score", "fscl_perm_round_commit",
           "fscl_perm_end", "fscl_bench_scan", "fscl_get_stats", "fscl_reset_stats", "fscl_fp64_peak", "fscl_device_gen_blocks",
           "fscl_host_alpha_tables", "fscl_host_enum_cells", "fscl_host_gen_blocks", "fscl_host_replay_round",
           "fscl_host_rng_draws", "fscl_project_pvalues", "fscl_ncx2_sf", "fscl_build_tables", "fscl_set_error",
           "fscl_set_exact_candidates", "fscl_perm_snapshot", "fscl_perm_seek", "fscl_debug_counters", "fscl_host_piece_maxima",
           "fscl_perm_fetch_clr", "fscl_perm_project", "fscl_comm_unique_id", "fscl_comm_init_rank", "fscl_comm_init_all",
           "fscl_comm_destroy", "fscl_comm_info", "fscl_comm_move", "fscl_debug_histogram", "fscl_host_replay_gathered",
           "fscl_set_endpoint_sharing"]

_lib = None


def load_library(path=LIB_PATH):
    """dlopen the CUDA library; raises (never falls back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise FsclError(f"{path} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(fscl_b200 has no CPU fallback)")
    L = C.CDLL(path)
    vp, ci = C.c_void_p, C.c_int
    L.fscl_last_error.restype = C.c_char_p
    L.fscl_create.argtypes = [vp, vp]
    L.fscl_destroy.argtypes = [vp]
    L.fscl_destroy.restype = None
    L.fscl_enum_cells.argtypes = [vp, ci, vp, vp, vp, ci]
    L.fscl_grid_eval.argtypes = [vp, ci, vp, vp, vp]
    L.fscl_search_maxpos.argtypes = [vp, ci, vp, vp, vp, vp]
    L.fscl_scan_chromosome.argtypes = [vp, ci, vp, ci, vp]
    L.fscl_block_permute.argtypes = [vp, vp, ci, vp, vp, ci]
    L.fscl_scan_permute.argtypes = [vp, vp, ci, vp, vp, vp]
    L.fscl_perm_begin.argtypes = [vp, vp, ci, vp, vp]
    L.fscl_perm_round_score.argtypes = [vp, vp, C.c_int64, vp, vp]
    L.fscl_perm_round_commit.argtypes = [vp, vp]
    L.fscl_perm_end.argtypes = [vp, vp]
    L.fscl_perm_snapshot.argtypes = [vp, vp]
    L.fscl_perm_seek.argtypes = [vp, ci]
    L.fscl_bench_scan.argtypes = [vp, ci, ci, ci, vp, vp]
    L.fscl_get_stats.argtypes = [vp, vp]
    L.fscl_set_exact_candidates.argtypes = [vp, ci]
    L.fscl_set_endpoint_sharing.argtypes = [vp, ci]
    L.fscl_debug_counters.argtypes = [vp, vp]
    L.fscl_debug_histogram.argtypes = [vp, vp]
    L.fscl_reset_stats.argtypes = [vp]
    L.fscl_fp64_peak.argtypes = [ci, vp]
    L.fscl_device_gen_blocks.argtypes = [vp, vp, ci, vp, ci, vp]
    L.fscl_host_alpha_tables.argtypes = [vp, vp, vp, vp]
    L.fscl_host_piece_maxima.argtypes = [ci, ci, C.c_double, vp, vp, vp, C.c_int64, vp]
    L.fscl_host_enum_cells.argtypes = [ci, vp, vp, ci, vp, vp, vp, ci]
    L.fscl_host_gen_blocks.argtypes = [ci, vp, vp, vp, ci, vp, ci, vp]
    L.fscl_host_replay_round.argtypes = [vp, vp, ci, vp, ci, ci, vp, vp]
    L.fscl_host_replay_gathered.argtypes = [vp, vp, ci, vp, ci, C.c_int64, ci, ci, vp, vp]
    L.fscl_host_rng_draws.argtypes = [ci, C.c_uint64, C.c_uint32, C.c_uint32, ci, vp]
    L.fscl_project_pvalues.argtypes = [ci, vp, vp, C.c_double, vp, vp]
    L.fscl_ncx2_sf.argtypes = [ci, vp, C.c_double, vp, vp]
    L.fscl_build_tables.argtypes = [ci, ci, vp, vp, ci, ci, ci, ci, ci, vp, vp, vp]
    L.fscl_perm_fetch_clr.argtypes = [vp, vp, vp, C.c_int64]
    L.fscl_perm_project.argtypes = [vp, C.c_double, vp, vp]
    L.fscl_comm_unique_id.argtypes = [vp]
    L.fscl_comm_init_rank.argtypes = [vp, ci, ci, vp]
    L.fscl_comm_init_all.argtypes = [vp, ci]
    L.fscl_comm_destroy.argtypes = [vp]
    L.fscl_comm_info.argtypes = [vp, vp, vp]
    L.fscl_comm_move.argtypes = [vp, vp]
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise FsclError(f"fscl_b200 error {rc}: {_lib.fscl_last_error().decode()}")


def comm_unique_id():
    """128 opaque bytes naming a new communicator (ncclGetUniqueId); make on one rank, ship to all."""
    L = load_library()
    buf = (C.c_char * 128)()
    _check(L.fscl_comm_unique_id(buf))
    return bytes(buf.raw)


def fp64_peak_tflops(device=0):
    L = load_library()
    v = C.c_double(0)
    _check(L.fscl_fp64_peak(device, C.byref(v)))
    return v.value


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


# ---- host-side pieces of the path (no device needed) ---------------------------
def host_alpha_tables(log_table=None):
    """(la[12][15], dmax[12][15], n[12]); row 11 is the coarse list (sm-search.c:276-296)."""
    L = load_library()
    la = np.zeros((12, 15)); dmax = np.zeros((12, 15), dtype=np.int32); n = np.zeros(12, dtype=np.int32)
    lt = None if log_table is None else np.ascontiguousarray(log_table, dtype=np.float64)
    _check(L.fscl_host_alpha_tables(None if lt is None else lt.ctypes.data, la.ctypes.data, dmax.ctypes.data,
                                    n.ctypes.data))
    return la, dmax, n


def host_piece_maxima(coef, null_tab, step):
    """[levels][n_knots+2][n_cls] upper bounds of (spline - null) over knot ranges (fscl_b200.h)."""
    L = load_library()
    coef = np.ascontiguousarray(coef, dtype=np.float64)
    null_tab = np.ascontiguousarray(null_tab, dtype=np.float64)
    n_cls, K = coef.shape[0], coef.shape[1]
    out = np.zeros(12 * (K + 2) * n_cls, dtype=np.float64)
    lv = C.c_int32(0)
    _check(L.fscl_host_piece_maxima(n_cls, K, float(step), coef.ctypes.data, null_tab.ctypes.data, out.ctypes.data,
                                    out.size, C.byref(lv)))
    return out[:lv.value * (K + 2) * n_cls].reshape(lv.value, K + 2, n_cls)


def host_enum_cells(chr_startpos, chr_bplen, G):
    L = load_library()
    sp, bl = _i32(chr_startpos), _i32(chr_bplen)
    n = L.fscl_host_enum_cells(len(sp), sp.ctypes.data, bl.ctypes.data, G, None, None, None, 0)
    if n < 0:
        _check(2)
    cc = np.zeros((3, max(n, 1)), dtype=np.int32)
    L.fscl_host_enum_cells(len(sp), sp.ctypes.data, bl.ctypes.data, G, cc[0].ctypes.data, cc[1].ctypes.data,
                           cc[2].ctypes.data, n)
    return cc[:, :n]


def host_gen_blocks(chr_, pos, t, nbp=0.1, width_mb=1.0, rng=RNG_PHILOX, seed=DEFAULT_SEED, max_blocks=1 << 16):
    """(i, j, len) triples of permutation t (draw phase of snp_block_permute) and the D7 count."""
    L = load_library()
    chr_, pos = _i32(chr_), _i32(pos)
    o = fscl_permute_opts_t(0, nbp, 1, width_mb, rng, seed, 1, 0, 1, 0, 0)
    blocks = np.zeros((max_blocks, 3), dtype=np.int32)
    d7 = C.c_int32(0)
    nb = L.fscl_host_gen_blocks(len(pos), chr_.ctypes.data, pos.ctypes.data, C.byref(o), t, blocks.ctypes.data,
                                max_blocks, C.byref(d7))
    if nb < 0:
        _check(2)
    return blocks[:min(nb, max_blocks)], d7.value


def host_replay_round(pts, active, matrix, t_first, seed=DEFAULT_SEED, prune=1, permute_clr=None):
    """Counting + stochastic pruning of one round; pts is updated in place."""
    L = load_library()
    active = _i32(active)
    matrix = np.ascontiguousarray(matrix, dtype=np.float64)
    o = fscl_permute_opts_t(0, 0.1, 1, 1.0, RNG_PHILOX, seed, prune, 0, 1, 0, 0)
    r = L.fscl_host_replay_round(pts.ctypes.data, active.ctypes.data, len(active), matrix.ctypes.data,
                                 matrix.shape[0], t_first, C.byref(o),
                                 None if permute_clr is None else permute_clr.ctypes.data)
    if r < 0:
        _check(2)
    return r


def host_replay_gathered(pts, active, gathered, world, chunk, R, t_first, seed=DEFAULT_SEED, prune=1, permute_clr=None):
    """Counting + pruning of one round from all-gathered rows (rank g's rows [q][n_active] at g * chunk)."""
    L = load_library()
    active = _i32(active)
    gathered = np.ascontiguousarray(gathered, dtype=np.float64)
    o = fscl_permute_opts_t(0, 0.1, 1, 1.0, RNG_PHILOX, seed, prune, 0, 1, 0, 0)
    r = L.fscl_host_replay_gathered(pts.ctypes.data, active.ctypes.data, len(active), gathered.ctypes.data, world, chunk,
                                    R, t_first, C.byref(o), None if permute_clr is None else permute_clr.ctypes.data)
    if r < 0:
        _check(2)
    return r


def build_tables(depths, fsp, spline_pts=200, asc_depth=0, asc_min_freq=1, ascbias_background_only=0,
                 include_invariant=0, device=0):
    """Sweep-model spline tables + null model on the GPU (compute_sweep_model_tables, sm-spline.c:486).
    fsp: list of background spectra, one per depth.  Returns (coef[n_cls][K][4], null_tab[n_cls], flags)."""
    L = load_library()
    depths = _i32(depths)
    fsp = [np.ascontiguousarray(f, dtype=np.float64) for f in fsp]
    n_cls = int(sum((d + 1) + (d // 2 + 1) for d in depths))
    coef = np.zeros((n_cls, spline_pts, 4))
    null = np.zeros(n_cls)
    ptrs = (C.c_void_p * len(fsp))(*[f.ctypes.data for f in fsp])
    flags = C.c_int32(0)
    _check(L.fscl_build_tables(device, len(depths), depths.ctypes.data, ptrs, spline_pts, asc_depth, asc_min_freq,
                               ascbias_background_only, include_invariant, coef.ctypes.data, null.ctypes.data,
                               C.byref(flags)))
    return coef, null, flags.value


def ncx2_sf(x, df, lam):
    L = load_library()
    x = np.ascontiguousarray(np.atleast_1d(x), dtype=np.float64)
    lam = np.ascontiguousarray(np.broadcast_to(lam, x.shape), dtype=np.float64)
    out = np.zeros_like(x)
    _check(L.fscl_ncx2_sf(len(x), x.ctypes.data, df, lam.ctypes.data, out.ctypes.data))
    return out


def project_pvalues(pts, permute_clr, df=2.0):
    """Projected p-values (reference README.md:11; defined in DESIGN.md 3.6): (lambda_hat, p)."""
    L = load_library()
    pts = np.ascontiguousarray(pts)
    clr = np.ascontiguousarray(permute_clr, dtype=np.float32)
    assert clr.shape == (len(pts), NULL_SAVE)
    lam, p = np.zeros(len(pts)), np.zeros(len(pts))
    _check(L.fscl_project_pvalues(len(pts), pts.ctypes.data, clr.ctypes.data, df, lam.ctypes.data, p.ctypes.data))
    return lam, p


def host_rng_draws(rng, seed, n, t=0, stream=0):
    L = load_library()
    out = np.zeros(n, dtype=np.int32)
    _check(L.fscl_host_rng_draws(rng, seed, t, stream, n, out.ctypes.data))
    return out


class Scanner:
    """One problem resident on one GPU.  ``flat`` holds the flattened arrays the
    C-ABI takes (see include/fscl_b200.h ``fscl_problem_t``)."""

    def __init__(self, flat, device=0, eval_range=81920, bp_resl=128, log_table=None):
        self.L = load_library()
        self._ctx = C.c_void_p(None)
        k = self._keep = {n: _i32(flat[n]) for n in
                          ("chr", "pos", "cls", "chr_start", "chr_nsnps", "chr_startpos", "chr_bplen")}
        k["coef"] = np.ascontiguousarray(flat["coef"], dtype=np.float64)
        k["null_tab"] = np.ascontiguousarray(flat["null_tab"], dtype=np.float64)
        if log_table is not None:
            k["log_table"] = np.ascontiguousarray(log_table, dtype=np.float64)
            assert k["log_table"].shape == (65536,)
        self.chr_names = list(flat.get("chr_names", [str(i) for i in range(len(k["chr_start"]))]))
        self.n_snps = len(k["pos"])
        p = fscl_problem_t(self.n_snps, k["chr"].ctypes.data, k["pos"].ctypes.data, k["cls"].ctypes.data,
                           len(k["chr_start"]), k["chr_start"].ctypes.data, k["chr_nsnps"].ctypes.data,
                           k["chr_startpos"].ctypes.data, k["chr_bplen"].ctypes.data, len(k["null_tab"]),
                           int(flat["n_knots"]), float(flat["step"]), k["coef"].ctypes.data,
                           k["null_tab"].ctypes.data,
                           k["log_table"].ctypes.data if log_table is not None else None,
                           eval_range, bp_resl, device, 0)
        _check(self.L.fscl_create(C.byref(p), C.byref(self._ctx)))
        self.device = device

    def close(self):
        if self._ctx:
            self.L.fscl_destroy(self._ctx)
            self._ctx = C.c_void_p(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- forward scan ---------------------------------------------------
    def enum_cells(self, G):
        n = self.L.fscl_enum_cells(self._ctx, G, None, None, None, 0)
        if n < 0:
            _check(2)
        cc = np.zeros((3, max(n, 1)), dtype=np.int32)
        self.L.fscl_enum_cells(self._ctx, G, cc[0].ctypes.data, cc[1].ctypes.data, cc[2].ctypes.data, n)
        return cc[:, :n]

    def grid_eval(self, chr_, pos):
        chr_, pos = _i32(np.atleast_1d(chr_)), _i32(np.atleast_1d(pos))
        out = np.zeros(len(pos), dtype=PT_DTYPE)
        _check(self.L.fscl_grid_eval(self._ctx, len(pos), chr_.ctypes.data, pos.ctypes.data, out.ctypes.data))
        return out

    def search_maxpos(self, chr_, start, end):
        chr_, start, end = _i32(np.atleast_1d(chr_)), _i32(np.atleast_1d(start)), _i32(np.atleast_1d(end))
        out = np.zeros(len(start), dtype=PT_DTYPE)
        _check(self.L.fscl_search_maxpos(self._ctx, len(start), chr_.ctypes.data, start.ctypes.data,
                                         end.ctypes.data, out.ctypes.data))
        return out

    def scan_chromosome(self, G):
        n = self.L.fscl_enum_cells(self._ctx, G, None, None, None, 0)
        out = np.zeros(max(n, 1), dtype=PT_DTYPE)
        n_out = C.c_int(0)
        _check(self.L.fscl_scan_chromosome(self._ctx, G, out.ctypes.data, len(out), C.byref(n_out)))
        return out[:n_out.value]

    # ---- permutations -----------------------------------------------------
    @staticmethod
    def permute_opts(n_permute, G, nbp=0.1, width_mb=1.0, rng=RNG_PHILOX, seed=DEFAULT_SEED, prune=1,
                     first_perm=0, perm_stride=1, round_perms=0, keep_clr=0):
        return fscl_permute_opts_t(n_permute, nbp, G, width_mb, rng, seed, prune, first_perm, perm_stride,
                                   round_perms, keep_clr)

    def block_permute(self, t, nbp=0.1, width_mb=1.0, rng=RNG_PHILOX, seed=DEFAULT_SEED, max_blocks=1 << 16):
        o = self.permute_opts(0, 1, nbp, width_mb, rng, seed)
        cls = np.zeros(self.n_snps, dtype=np.int32)
        blocks = np.zeros((max_blocks, 3), dtype=np.int32)
        nb = self.L.fscl_block_permute(self._ctx, C.byref(o), t, cls.ctypes.data, blocks.ctypes.data, max_blocks)
        if nb < 0:
            _check(1)
        return cls, blocks[:min(nb, max_blocks)]

    def device_gen_blocks(self, t, nbp=0.1, width_mb=1.0, seed=DEFAULT_SEED, max_blocks=1 << 16):
        """(i, j, len) triples of permutation t drawn by the device kernel, and its flags."""
        o = self.permute_opts(0, 1, nbp, width_mb, RNG_PHILOX, seed)
        blocks = np.zeros((max_blocks, 3), dtype=np.int32)
        flags = C.c_int32(0)
        nb = self.L.fscl_device_gen_blocks(self._ctx, C.byref(o), t, blocks.ctypes.data, max_blocks, C.byref(flags))
        if nb < 0:
            _check(1)
        return blocks[:min(nb, max_blocks)], flags.value

    def scan_permute(self, pts, n_permute, G, save_clr=False, **kw):
        pts = np.ascontiguousarray(pts.copy())
        o = self.permute_opts(n_permute, G, **kw)
        clr = np.zeros((len(pts), NULL_SAVE), dtype=np.float32) if save_clr else None
        st = fscl_stats_t()
        _check(self.L.fscl_scan_permute(self._ctx, pts.ctypes.data, len(pts), C.byref(o),
                                        clr.ctypes.data if save_clr else None, C.byref(st)))
        return pts, st.asdict(), clr

    def scan_permute_rounds(self, pts, n_permute, G, save_clr=False, start_t=0, clr=None, max_rounds=None, **kw):
        """The permutation test round by round, interruptible and resumable (counter-based stream).

        Starts at permutation ``start_t`` with the counters found in ``pts`` (and the stored CLRs in
        ``clr``), runs at most ``max_rounds`` rounds and returns ``(pts, stats, clr, t_next)``:
        feeding these back as ``pts, clr, start_t=t_next`` continues the test, and the final result
        equals an uninterrupted ``scan_permute`` bit for bit.  ``t_next > n_permute`` means finished."""
        pts = np.ascontiguousarray(pts.copy())
        o = self.permute_opts(n_permute, G, **kw)
        if save_clr and clr is None:
            clr = np.zeros((len(pts), NULL_SAVE), dtype=np.float32)
        elif clr is not None:
            clr = np.ascontiguousarray(clr.copy())
        _check(self.L.fscl_perm_begin(self._ctx, pts.ctypes.data, len(pts), C.byref(o),
                                      clr.ctypes.data if clr is not None else None))
        if start_t:
            _check(self.L.fscl_perm_seek(self._ctx, int(start_t)))
        R, na, rounds = C.c_int(0), C.c_int(0), 0
        while max_rounds is None or rounds < max_rounds:
            _check(self.L.fscl_perm_round_score(self._ctx, None, 0, C.byref(R), C.byref(na)))
            if R.value == 0:
                break
            _check(self.L.fscl_perm_round_commit(self._ctx, None))
            rounds += 1
        t_next = C.c_int32(0)
        _check(self.L.fscl_perm_snapshot(self._ctx, C.byref(t_next)))
        st = fscl_stats_t()
        _check(self.L.fscl_perm_end(self._ctx, C.byref(st)))
        return pts, st.asdict(), clr, t_next.value

    def scan_permute_sharded(self, pts, n_permute, G, rank, world, allreduce, matrix, matrix_ptr, save_clr=False,
                             **kw):
        """Permutations sharded over ``world`` ranks (one process per GPU).

        ``matrix`` is a device buffer of float64 owned by the caller (for example
        a torch tensor), ``matrix_ptr`` its device address, and
        ``allreduce(matrix, n_elems)`` sums its first n_elems over the ranks
        (NCCL over NVLink).  Every rank ends with identical counts."""
        pts = np.ascontiguousarray(pts.copy())
        o = self.permute_opts(n_permute, G, first_perm=rank, perm_stride=world, **kw)
        clr = np.zeros((len(pts), NULL_SAVE), dtype=np.float32) if save_clr else None
        _check(self.L.fscl_perm_begin(self._ctx, pts.ctypes.data, len(pts), C.byref(o),
                                      clr.ctypes.data if save_clr else None))
        R, na = C.c_int(0), C.c_int(0)
        cap = matrix.numel() if hasattr(matrix, "numel") else len(matrix)
        while True:
            _check(self.L.fscl_perm_round_score(self._ctx, matrix_ptr, cap, C.byref(R), C.byref(na)))
            if R.value == 0:
                break
            allreduce(matrix, R.value * na.value)
            _check(self.L.fscl_perm_round_commit(self._ctx, matrix_ptr))
        st = fscl_stats_t()
        _check(self.L.fscl_perm_end(self._ctx, C.byref(st)))
        return pts, st.asdict(), clr

    # ---- multi-GPU ----------------------------------------------------------
    def comm_init_rank(self, world, rank, unique_id):
        """Join a communicator of ``world`` contexts, one process per GPU.  ``unique_id`` = the 128 bytes
        of ``comm_unique_id()`` made on rank 0 and shipped to every rank.  From then on
        ``scan_chromosome`` shards over cells and ``scan_permute`` over permutations (NCCL all-gather per
        round inside the library); every rank must make the same calls and gets identical results."""
        buf = (C.c_char * 128).from_buffer_copy(bytes(unique_id))
        _check(self.L.fscl_comm_init_rank(self._ctx, world, rank, buf))

    def comm_take(self, other):
        """Adopt the communicator of another Scanner on the same device (it is left without one)."""
        _check(self.L.fscl_comm_move(self._ctx, other._ctx))

    def comm_info(self):
        w, r = C.c_int(1), C.c_int(0)
        _check(self.L.fscl_comm_info(self._ctx, C.byref(w), C.byref(r)))
        return w.value, r.value

    # ---- stored permuted CLRs / projection, device-resident -------------------
    def fetch_clr(self, n_pts):
        """(offsets[n_pts+1], packed float32): the stored permuted CLRs of the last test run with
        ``save_clr`` / ``keep_clr`` -- cell i is packed[offsets[i]:offsets[i+1]] (one gather + one copy)."""
        off = np.zeros(n_pts + 1, dtype=np.int64)
        _check(self.L.fscl_perm_fetch_clr(self._ctx, off.ctypes.data, None, 0))
        packed = np.zeros(max(int(off[-1]), 1), dtype=np.float32)
        _check(self.L.fscl_perm_fetch_clr(self._ctx, off.ctypes.data, packed.ctypes.data, packed.size))
        return off, packed[:int(off[-1])]

    def project(self, n_pts, df=2.0):
        """Projected p-values computed on the device from the CLRs kept there: (lambda_hat, p)."""
        lam, p = np.zeros(n_pts), np.zeros(n_pts)
        _check(self.L.fscl_perm_project(self._ctx, float(df), lam.ctypes.data, p.ctypes.data))
        return lam, p

    # ---- measurement --------------------------------------------------------
    def bench_scan(self, G, n_perm_slots, count_terms=False):
        terms = C.c_int64(0)
        st = fscl_stats_t()
        _check(self.L.fscl_bench_scan(self._ctx, G, n_perm_slots, int(count_terms), C.byref(terms), C.byref(st)))
        return terms.value, st.asdict()

    def set_exact_candidates(self, on=True):
        """on: evaluate every coarse alpha candidate site by site (no pruning by bounds); results are
        identical either way (fscl_b200.h), only the work differs."""
        _check(self.L.fscl_set_exact_candidates(self._ctx, 1 if on else 0))

    def set_endpoint_sharing(self, on=True):
        """off: every cell evaluates both of its end points itself, as search_maxpos does in the reference; on (default):
        a cell that starts where the previous one ended takes that evaluation over.  Results are identical bit for bit."""
        _check(self.L.fscl_set_endpoint_sharing(self._ctx, 1 if on else 0))

    def stats(self):
        st = fscl_stats_t()
        _check(self.L.fscl_get_stats(self._ctx, C.byref(st)))
        return st.asdict()

    def reset_stats(self):
        _check(self.L.fscl_reset_stats(self._ctx))

    # ---- output -------------------------------------------------------------
    def format_records(self, pts, n_permute=0, label=None, maximum_only=False):
        return format_records(pts, self.chr_names, self._keep["pos"], n_permute, label, maximum_only)


def _c_fmt_e(v):
    """printf("%1.3e")"""
    return "%1.3e" % v


def format_records(pts, chr_names, pos, n_permute=0, label=None, maximum_only=False):
    """scan_output records (scan-chromosome.c:707-746)."""
    pre = f"{label}\t" if label else ""
    out = []
    if maximum_only:
        best = 0
        for i in range(1, len(pts)):
            if pts[i]["clr"] > pts[best]["clr"]:
                best = i
        p = pts[best]
        return (f"{pre}{chr_names[p['chr']]}\t{p['sweep_pos']}\t{'%1.2f' % p['clr']}\t"
                f"{_c_fmt_e(math.exp(p['lalpha']))}\t{p['n_snps']}\t{pos[p['window_start']]}\t"
                f"{pos[p['window_end']]}\n")
    for p in pts:
        head = f"{pre}{chr_names[p['chr']]}\t{p['sweep_pos']}\t{'%1.2f' % p['clr']}\t{_c_fmt_e(math.exp(p['lalpha']))}"
        if n_permute > 0:
            if p["permute_p"] < 2:
                pv = 1.0 / p["permute_n"]
            else:
                pv = (p["permute_p"] - 1.0) / (p["permute_n"] - 1.0)
            out.append(f"{head}\t{p['permute_p']}\t{p['permute_n']}\t{'%1.3f' % (-math.log10(pv))}\n")
        else:
            out.append(f"{head}\t{p['n_snps']}\t{pos[p['window_start']]}\t{pos[p['window_end']]}\n")
    return "".join(out)


def format_nulldist(pts, chr_names, permute_clr):
    """`<out>-nulldist` companion file (scan-chromosome.c:753-796): header with the 10 000 quantile
    labels, then per cell the stored permuted CLRs sorted ascending as %1.2f."""
    head = "chr\tpos\tCLR\talpha\tp\tn" + "".join("\t%1.4f" % (j / float(NULL_SAVE)) for j in range(NULL_SAVE)) + "\n"
    out = [head]
    for i, p in enumerate(pts):
        m = min(int(p["permute_n"]), NULL_SAVE)
        vals = np.sort(np.asarray(permute_clr[i, :m], dtype=np.float32))
        out.append(f"{chr_names[p['chr']]}\t{p['sweep_pos']}\t{'%1.3f' % p['clr']}\t{'%1.3e' % math.exp(p['lalpha'])}"
                   f"\t{p['permute_p']}\t{p['permute_n']}" + "".join("\t%1.2f" % float(v) for v in vals) + "\n")
    return "".join(out)
