"""ctypes bindings for the test oracle -- TEST INFRASTRUCTURE ONLY.

Two libraries are bound here:

* ``oracle/libfscl_oracle.so``   our plain-C restatement (fscl_oracle.c)
* ``oracle/_ref/libfscl_ref.so`` the UNMODIFIED reference (slowkoni/fscl)
  compiled from /root/reference by ``oracle/Makefile`` (struct mirrors below
  follow /root/reference/fscl.h:7-76).

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` leg may import this module; nothing under ``fscl_b200/``
does.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libfscl_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libfscl_ref.so")
REF_EXE = os.path.join(HERE, "_ref", "fscl_ref")
NULL_SAVE = 10000


def build(quiet=True):
    """make -C oracle (restatement always; _ref only where /root/reference exists)."""
    subprocess.run(["make", "-C", HERE], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


# --------------------------------------------------------------------------
# reference struct mirrors (fscl.h)
# --------------------------------------------------------------------------
class snp_t(C.Structure):            # fscl.h:7-14
    _fields_ = [("chr", C.c_int), ("pos", C.c_int), ("null_logl", C.c_double),
                ("obs_freq", C.c_int), ("depth_p", C.c_int), ("folded", C.c_int)]


SNP_DTYPE = np.dtype({"names": ["chr", "pos", "null_logl", "obs_freq", "depth_p", "folded"],
                      "formats": ["<i4", "<i4", "<f8", "<i4", "<i4", "<i4"],
                      "offsets": [0, 4, 8, 16, 20, 24], "itemsize": 32})


class chr_limits_t(C.Structure):     # fscl.h:26-33
    _fields_ = [("chr", C.c_int), ("name", C.c_char_p), ("start_index", C.c_int),
                ("n_snps", C.c_int), ("start_pos", C.c_int), ("bp_length", C.c_int)]


class scan_pt_t(C.Structure):        # fscl.h:35-51
    _fields_ = [("chr", C.c_int), ("nearest_snp", C.c_int), ("sweep_pos", C.c_int),
                ("n_snps", C.c_int), ("window_start", C.c_int), ("window_end", C.c_int),
                ("lalpha", C.c_double), ("null_logl", C.c_double), ("sm_logl", C.c_double),
                ("clr", C.c_double), ("permute_n", C.c_int), ("permute_p", C.c_int),
                ("permute_finished", C.c_int), ("scan_running", C.c_int),
                ("permute_clr", C.POINTER(C.c_float))]


class scan_t(C.Structure):           # fscl.h:53-62
    _fields_ = [("n_snps", C.c_int), ("snps", C.POINTER(snp_t)), ("n_depths", C.c_int),
                ("sample_depths", C.POINTER(C.c_int)), ("n_scan_pts", C.c_int),
                ("scan_pts", C.POINTER(scan_pt_t)), ("chr_limits", C.POINTER(chr_limits_t)),
                ("n_chromosomes", C.c_int)]


class spline_t(C.Structure):         # fscl.h:64-68
    _fields_ = [("n", C.c_int), ("knot_points", C.POINTER(C.c_double)),
                ("coef", C.POINTER(C.POINTER(C.c_double)))]


class sm_ptable_t(C.Structure):      # fscl.h:70-76
    _fields_ = [("spline_func", C.POINTER(C.POINTER(spline_t))),
                ("fspline_func", C.POINTER(C.POINTER(spline_t))),
                ("sample_size", C.c_int), ("pbk", C.POINTER(C.POINTER(C.c_double))),
                ("fsp", C.POINTER(C.c_double))]


_ref = None
_libc = None


def libc():
    global _libc
    if _libc is None:
        _libc = C.CDLL(None)
        _libc.srand.argtypes = [C.c_uint]
        _libc.rand.restype = C.c_int
    return _libc


def _libm_log(x):
    """log() of the C library, as compute_snp_null_model calls it (numpy's vectorised log differs in the last
    ulp on some CPUs)."""
    x = float(x)
    if x > 0.0:
        return math.log(x)
    return -math.inf if x == 0.0 else math.nan


def ref_available():
    return os.path.exists(REF_SO)


def ref():
    """The unmodified reference as a shared library (kind 'reference')."""
    global _ref
    if _ref is None:
        L = C.CDLL(REF_SO, mode=C.RTLD_GLOBAL)
        L.load_snp_input.restype = C.POINTER(scan_t)
        L.load_snp_input.argtypes = [C.c_char_p, C.c_int, C.c_int]
        L.background_fsp.restype = C.POINTER(C.POINTER(C.c_double))
        L.background_fsp.argtypes = [C.POINTER(scan_t), C.c_int, C.c_char_p, C.c_int]
        L.compute_sweep_model_tables.restype = C.POINTER(sm_ptable_t)
        L.compute_sweep_model_tables.argtypes = [C.POINTER(scan_t), C.POINTER(C.POINTER(C.c_double)),
                                                 C.c_int, C.c_int, C.c_int, C.c_int]
        L.compute_snp_null_model.argtypes = [C.POINTER(scan_t), C.POINTER(C.POINTER(C.c_double))]
        L.scan_chromosome.argtypes = [C.POINTER(scan_t), C.POINTER(sm_ptable_t)] + [C.c_int] * 4
        L.scan_permute.argtypes = [C.POINTER(scan_t), C.POINTER(sm_ptable_t), C.c_int, C.c_double,
                                   C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double]
        L.scan_output.argtypes = [C.c_char_p, C.POINTER(scan_t), C.c_int, C.c_int, C.c_char_p]
        L.search_maxalpha.argtypes = [C.POINTER(scan_pt_t), C.POINTER(snp_t), C.POINTER(sm_ptable_t)]
        L.spline_interpolate.restype = C.c_double
        L.spline_interpolate.argtypes = [C.POINTER(spline_t), C.c_double]
        L.fx_snp_block_permute.argtypes = [C.POINTER(snp_t), C.POINTER(snp_t), C.c_int, C.c_double, C.c_double]
        L.fx_init_scan_result.argtypes = [C.POINTER(scan_pt_t), C.c_int, C.POINTER(snp_t),
                                          C.POINTER(chr_limits_t), C.c_int, C.c_int]
        L.fx_search_maxpos.argtypes = [C.POINTER(scan_pt_t), C.c_int, C.c_int, C.c_int, C.POINTER(snp_t),
                                       C.POINTER(chr_limits_t), C.c_int, C.c_int, C.POINTER(sm_ptable_t)]
        L.lchoose.restype = C.c_double
        L.lchoose.argtypes = [C.c_int, C.c_int]
        L.configure_logmsg.argtypes = [C.c_int]
        L.configure_logmsg(1)
        C.c_int.in_dll(L, "spline_pts").value = 200
        L.init_log_table()
        os.environ.setdefault("OMP_NUM_THREADS", "1")          # D9
        _ref = L
    return _ref


def write_snp_table(table, path):
    names = table["chr_names"]
    c, p, k, d, f = (table[x] for x in ("chr", "pos", "count", "depth", "folded"))
    with open(path, "w") as fh:
        fh.writelines(f"{names[c[i]]}\t{p[i]}\t{k[i]}\t{d[i]}\t{f[i]}\n" for i in range(len(p)))


class RefProblem:
    """A problem instance prepared by the reference's own host code:
    load_snp_input -> background_fsp -> compute_sweep_model_tables ->
    compute_snp_null_model (fscl.c:316-326)."""

    def __init__(self, table, asc_depth=0, asc_min_freq=1, ascbias_background_only=0,
                 include_invariant=0, force_neutral=0, minimum_depth=5, spline_pts=200):
        L = ref()
        C.c_int.in_dll(L, "spline_pts").value = spline_pts
        fd, path = tempfile.mkstemp(suffix=".snps")
        os.close(fd)
        try:
            write_snp_table(table, path)
            devnull = os.open(os.devnull, os.O_WRONLY)
            saved = os.dup(2)
            os.dup2(devnull, 2)
            try:
                self.scan = L.load_snp_input(path.encode(), include_invariant, minimum_depth)
            finally:
                os.dup2(saved, 2)
                os.close(saved)
                os.close(devnull)
        finally:
            os.unlink(path)
        self.L = L
        self.fsp = L.background_fsp(self.scan, force_neutral, None, include_invariant)
        self.sm_p = L.compute_sweep_model_tables(self.scan, self.fsp, asc_depth, asc_min_freq,
                                                 ascbias_background_only, include_invariant)
        L.compute_snp_null_model(self.scan, self.fsp)
        self.spline_pts = spline_pts

    # -- flatten into the arrays the oracle / the C-ABI take ---------------
    def snps(self):
        s = self.scan.contents
        buf = (C.c_char * (32 * s.n_snps)).from_address(C.addressof(s.snps.contents))
        return np.frombuffer(buf, dtype=SNP_DTYPE)

    def flatten(self):
        s = self.scan.contents
        snps = self.snps()
        depths = np.array([s.sample_depths[i] for i in range(s.n_depths)], dtype=np.int32)
        base = np.zeros(s.n_depths + 1, dtype=np.int64)
        for d in range(s.n_depths):
            base[d + 1] = base[d] + (depths[d] + 1) + (depths[d] // 2 + 1)
        n_cls = int(base[-1])
        K = self.spline_pts
        coef = np.zeros((n_cls, K, 4), dtype=np.float64)
        null_tab = np.zeros(n_cls, dtype=np.float64)
        for d in range(s.n_depths):
            n = int(depths[d])
            fsp = np.array([self.fsp[d][f] for f in range(n + 1)])
            t = self.sm_p[d]
            for f in range(n + 1):
                sp = t.spline_func[f].contents
                coef[base[d] + f] = np.ctypeslib.as_array(sp.coef[0], shape=(K, 4))
                null_tab[base[d] + f] = _libm_log(fsp[f])
            for f in range(n // 2 + 1):
                sp = t.fspline_func[f].contents
                coef[base[d] + n + 1 + f] = np.ctypeslib.as_array(sp.coef[0], shape=(K, 4))
                null_tab[base[d] + n + 1 + f] = _libm_log(fsp[f] + fsp[n - f]) if f != n - f else _libm_log(fsp[f])
        cls = np.where(snps["folded"] != 0,
                       base[snps["depth_p"]] + depths[snps["depth_p"]] + 1 + snps["obs_freq"],
                       base[snps["depth_p"]] + snps["obs_freq"]).astype(np.int32)
        lim = [s.chr_limits[i] for i in range(s.n_chromosomes)]
        flat = dict(
            chr=np.ascontiguousarray(snps["chr"]), pos=np.ascontiguousarray(snps["pos"]), cls=cls,
            chr_names=[l.name.decode() for l in lim],
            chr_start=np.array([l.start_index for l in lim], dtype=np.int32),
            chr_nsnps=np.array([l.n_snps for l in lim], dtype=np.int32),
            chr_startpos=np.array([l.start_pos for l in lim], dtype=np.int32),
            chr_bplen=np.array([l.bp_length for l in lim], dtype=np.int32),
            coef=coef, null_tab=null_tab, n_knots=K, step=24.0 / (K + 1.0),
            depths=depths, cls_base=base,
        )
        # the reference's per-site null must be the class table's value bit for bit
        assert np.array_equal(null_tab[cls], snps["null_logl"])
        return flat


# --------------------------------------------------------------------------
# restatement bindings
# --------------------------------------------------------------------------
class fo_ctx_t(C.Structure):
    _fields_ = [("n_snps", C.c_int), ("chr", C.c_void_p), ("pos", C.c_void_p), ("n_chr", C.c_int),
                ("chr_start", C.c_void_p), ("chr_nsnps", C.c_void_p), ("chr_startpos", C.c_void_p),
                ("chr_bplen", C.c_void_p), ("n_cls", C.c_int), ("n_knots", C.c_int), ("step", C.c_double),
                ("coef", C.c_void_p), ("null_tab", C.c_void_p), ("logT", C.c_void_p),
                ("eval_range", C.c_int), ("bp_resl", C.c_int), ("mode", C.c_int)]


class fo_pt_t(C.Structure):
    _fields_ = [("chr", C.c_int), ("nearest", C.c_int), ("sweep_pos", C.c_int), ("n_snps", C.c_int),
                ("wstart", C.c_int), ("wend", C.c_int), ("lalpha", C.c_double), ("null_logl", C.c_double),
                ("sm_logl", C.c_double), ("clr", C.c_double), ("permute_n", C.c_int),
                ("permute_p", C.c_int), ("finished", C.c_int), ("pad_", C.c_int)]


PT_DTYPE = np.dtype([("chr", "<i4"), ("nearest", "<i4"), ("sweep_pos", "<i4"), ("n_snps", "<i4"),
                     ("wstart", "<i4"), ("wend", "<i4"), ("lalpha", "<f8"), ("null_logl", "<f8"),
                     ("sm_logl", "<f8"), ("clr", "<f8"), ("permute_n", "<i4"), ("permute_p", "<i4"),
                     ("finished", "<i4"), ("pad_", "<i4")])
assert PT_DTYPE.itemsize == C.sizeof(fo_pt_t) == 72


class fo_rng_t(C.Structure):
    _fields_ = [("kind", C.c_int), ("r", C.c_int32 * 34), ("k", C.c_int), ("key0", C.c_uint32),
                ("key1", C.c_uint32), ("perm", C.c_uint32), ("ctr", C.c_uint32), ("n_draws", C.c_uint64)]


_orc = None


def lib():
    global _orc
    if _orc is None:
        if not os.path.exists(ORACLE_SO):
            build()
        L = C.CDLL(ORACLE_SO)
        vp, ci, cd = C.c_void_p, C.c_int, C.c_double
        L.fo_init_log_table.argtypes = [vp]
        L.fo_logt.restype = cd
        L.fo_logt.argtypes = [vp, ci]
        L.fo_dmax.argtypes = [vp, cd]
        L.fo_spline_eval.restype = cd
        L.fo_spline_eval.argtypes = [vp, ci, cd, cd]
        L.fo_alpha_lists.argtypes = [vp, ci, vp]
        L.fo_search_snppos.argtypes = [vp, ci, cd]
        L.fo_init_scan_result.argtypes = [vp, vp, vp, ci, ci]
        L.fo_search_maxalpha.argtypes = [vp, vp, vp]
        L.fo_search_maxpos.argtypes = [vp, vp, ci, ci, ci, vp, vp]
        L.fo_enum_cells.argtypes = [vp, ci, vp, vp, vp]
        L.fo_count_cells.argtypes = [vp, ci]
        L.fo_scan_cells.argtypes = [vp, vp, ci, vp, vp]
        L.fo_count_terms_maxalpha.restype = C.c_long
        L.fo_count_terms_maxalpha.argtypes = [vp, vp, vp]
        L.fo_rng_seed_glibc.argtypes = [vp, C.c_uint]
        L.fo_rng_seed_philox.argtypes = [vp, C.c_uint64]
        L.fo_rng_set_perm.argtypes = [vp, C.c_uint32]
        L.fo_rng_next.argtypes = [vp]
        L.fo_rng_next.restype = C.c_int32
        L.fo_philox_draw.restype = C.c_int32
        L.fo_philox_draw.argtypes = [C.c_uint32] * 5
        L.fo_philox4x32_10.argtypes = [vp, vp, vp]
        L.fo_block_permute.argtypes = [vp, vp, cd, cd, vp, vp, vp, ci, vp]
        L.fo_scan_permute.argtypes = [vp, vp, vp, ci, ci, cd, ci, cd, vp, ci, vp, vp, vp]
        L.fo_format_record.argtypes = [C.c_char_p, ci, C.c_char_p, C.c_char_p, vp, ci, ci, ci]
        L.fo_lchoose.restype = cd
        L.fo_lchoose.argtypes = [ci, ci]
        L.fo_estimate_spline.argtypes = [vp, vp, ci, vp]
        L.fo_sweep_model_tables.argtypes = [vp, ci, ci, ci, ci, ci, ci, vp, vp]
        L.fo_ascbias_adjust_background.argtypes = [vp, ci, ci, ci, vp]
        L.fo_background_fsp.argtypes = [ci, vp, vp, vp, ci, vp, ci, ci, vp]
        L.fo_null_model.argtypes = [vp, ci, vp, vp]
        _orc = L
    return _orc


_logT = None


def log_table():
    global _logT
    if _logT is None:
        _logT = np.zeros(65536, dtype=np.float64)
        lib().fo_init_log_table(_logT.ctypes.data)
    return _logT


def _p(a):
    return a.ctypes.data


class Oracle:
    """The restatement bound to one flattened problem (see RefProblem.flatten)."""

    def __init__(self, flat, mode=1, eval_range=81920, bp_resl=128):
        self.L = lib()
        self.flat = flat
        self.keep = {k: np.ascontiguousarray(flat[k]) for k in
                     ("chr", "pos", "cls", "chr_start", "chr_nsnps", "chr_startpos", "chr_bplen", "coef", "null_tab")}
        k = self.keep
        assert k["chr"].dtype == np.int32 and k["pos"].dtype == np.int32 and k["cls"].dtype == np.int32
        self.logT = log_table()
        self.ctx = fo_ctx_t(len(k["pos"]), _p(k["chr"]), _p(k["pos"]), len(k["chr_start"]), _p(k["chr_start"]),
                            _p(k["chr_nsnps"]), _p(k["chr_startpos"]), _p(k["chr_bplen"]), len(k["null_tab"]),
                            flat["n_knots"], flat["step"], _p(k["coef"]), _p(k["null_tab"]), _p(self.logT),
                            eval_range, bp_resl, mode)
        self.cls = k["cls"]

    @property
    def n_snps(self):
        return self.ctx.n_snps

    def ctxp(self):
        return C.addressof(self.ctx)

    def grid_eval(self, chr_, pos, cls=None):
        cls = self.cls if cls is None else cls
        pt = np.zeros(1, dtype=PT_DTYPE)
        self.L.fo_init_scan_result(self.ctxp(), _p(cls), _p(pt), chr_, pos)
        self.L.fo_search_maxalpha(self.ctxp(), _p(cls), _p(pt))
        return pt[0]

    def enum_cells(self, G):
        n = self.L.fo_count_cells(self.ctxp(), G)
        cc = np.zeros((3, n), dtype=np.int32)
        self.L.fo_enum_cells(self.ctxp(), G, _p(cc[0]), _p(cc[1]), _p(cc[2]))
        return cc

    def search_maxpos(self, chr_, start, end, cls=None):
        cls = self.cls if cls is None else cls
        pt = np.zeros(1, dtype=PT_DTYPE)
        ne = C.c_long(0)
        self.L.fo_search_maxpos(self.ctxp(), _p(cls), chr_, start, end, _p(pt), C.addressof(ne))
        return pt[0], ne.value

    def scan(self, G, cls=None):
        cls = self.cls if cls is None else cls
        n = self.L.fo_count_cells(self.ctxp(), G)
        pts = np.zeros(n, dtype=PT_DTYPE)
        ne = C.c_long(0)
        self.L.fo_scan_cells(self.ctxp(), _p(cls), G, _p(pts), C.addressof(ne))
        return pts, ne.value

    def count_terms(self, pts, cls=None):
        cls = self.cls if cls is None else cls
        return sum(self.L.fo_count_terms_maxalpha(self.ctxp(), _p(cls), _p(pts[i:i + 1])) for i in range(len(pts)))

    def rng(self, kind, seed):
        g = fo_rng_t()
        if kind == "glibc":
            self.L.fo_rng_seed_glibc(C.addressof(g), seed)
        else:
            self.L.fo_rng_seed_philox(C.addressof(g), seed)
        return g

    def block_permute(self, g, nbp=0.1, width_mb=1.0, cls=None, max_blocks=1 << 16):
        cls = self.cls if cls is None else cls
        out = np.empty_like(cls)
        blocks = np.zeros((max_blocks, 3), dtype=np.int32)
        flags = C.c_int(0)
        nb = self.L.fo_block_permute(self.ctxp(), C.addressof(g), nbp, width_mb, _p(cls), _p(out), _p(blocks),
                                     max_blocks, C.addressof(flags))
        return out, blocks[:min(nb, max_blocks)], flags.value

    def scan_permute(self, pts, n_permute, G, g, nbp=0.1, width_mb=1.0, prune=1, save_clr=False):
        pts = pts.copy()
        clr = np.zeros((len(pts), NULL_SAVE), dtype=np.float32) if save_clr else None
        flags = C.c_int(0)
        ne = C.c_long(0)
        scored = self.L.fo_scan_permute(self.ctxp(), _p(self.cls), _p(pts), len(pts), n_permute, nbp, G, width_mb,
                                        C.addressof(g), prune, _p(clr) if save_clr else None,
                                        C.addressof(flags), C.addressof(ne))
        return pts, scored, flags.value, ne.value, clr

    def format_records(self, pts, n_permute=0, label=None):
        buf = C.create_string_buffer(4096)
        names, pos = self.flat["chr_names"], self.keep["pos"]
        out = []
        for i in range(len(pts)):
            n = self.L.fo_format_record(buf, 4096, label.encode() if label else None,
                                        names[pts[i]["chr"]].encode(), _p(pts[i:i + 1]), n_permute,
                                        int(pos[pts[i]["wstart"]]), int(pos[pts[i]["wend"]]))
            out.append(buf.raw[:n].decode())
        return "".join(out)


def run_ref_exe(table, out_path, G=100000, n_permute=0, extra=(), threads=1, quiet=True):
    """Run the unmodified reference executable on a synthetic table (kind 'reference')."""
    fd, path = tempfile.mkstemp(suffix=".snps")
    os.close(fd)
    try:
        write_snp_table(table, path)
        cmd = [REF_EXE, "-f", path, "-o", out_path, "-G", str(G), "-p", str(n_permute),
               f"--n-threads={threads}", *extra]
        env = dict(os.environ, OMP_NUM_THREADS="1" if threads == 1 else str(threads))
        subprocess.run(cmd, check=True, env=env, stdout=subprocess.DEVNULL,
                       stderr=subprocess.DEVNULL if quiet else None)
    finally:
        os.unlink(path)
    with open(out_path) as fh:
        return fh.read()


# --------------------------------------------------------------------------
# oracle-only problem preparation (no reference library needed)
# --------------------------------------------------------------------------
def load_snp_table_text(text, include_invariant=0, minimum_depth=5):
    """Restatement of load_snp_input (snp-input.c:19-145) for well-formed input:
    filters, folding to the minor count, first-seen depth and chromosome
    indices, sort by (chromosome index, position)."""
    names, chr_, pos, cnt, dep, fol = [], [], [], [], [], []
    for line in text.splitlines():
        line = line.strip()
        if not line or line.startswith("#"):
            continue
        f = line.split()
        if len(f) < 5:
            continue
        try:
            p, k, d, fo = int(f[1]), int(f[2]), int(f[3]), int(f[4])
        except ValueError:
            continue
        if d < minimum_depth:
            continue
        if not include_invariant and (k < 1 or k > d - 1):
            continue
        if f[0] not in names:
            names.append(f[0])
        if fo and k > d - k:
            k = d - k
        chr_.append(names.index(f[0])); pos.append(p); cnt.append(k); dep.append(d); fol.append(fo)
    chr_, pos, cnt, dep, fol = (np.asarray(a, dtype=np.int32) for a in (chr_, pos, cnt, dep, fol))
    depths = []
    for d in dep:                       # first-seen order (snp-input.c:97-104), before the sort
        if d not in depths:
            depths.append(int(d))
    order = np.lexsort((pos, chr_))     # stable; parity inputs have unique (chr, pos)
    chr_, pos, cnt, dep, fol = (a[order] for a in (chr_, pos, cnt, dep, fol))
    depth_p = np.array([depths.index(int(d)) for d in dep], dtype=np.int32)
    return dict(chr_names=names, chr=chr_, pos=pos, obs_freq=cnt, depth_p=depth_p, folded=fol,
                depths=np.array(depths, dtype=np.int32))


def oracle_flatten(snp, asc_depth=0, asc_min_freq=1, ascbias_background_only=0, include_invariant=0,
                   force_neutral=0, spline_pts=200):
    """background_fsp -> compute_sweep_model_tables -> compute_snp_null_model with the
    restatement only; returns the same dict as RefProblem.flatten()."""
    L = lib()
    depths = snp["depths"]
    n_snps = len(snp["pos"])
    fsp = [np.zeros(int(d) + 1) for d in depths]
    ptrs = (C.c_void_p * len(depths))(*[f.ctypes.data for f in fsp])
    of, dp, fo = (np.ascontiguousarray(snp[k], dtype=np.int32) for k in ("obs_freq", "depth_p", "folded"))
    L.fo_background_fsp(n_snps, of.ctypes.data, dp.ctypes.data, fo.ctypes.data, len(depths), depths.ctypes.data,
                        force_neutral, include_invariant, ptrs)
    base = np.zeros(len(depths) + 1, dtype=np.int64)
    for d in range(len(depths)):
        base[d + 1] = base[d] + (depths[d] + 1) + (depths[d] // 2 + 1)
    K = spline_pts
    coef = np.zeros((int(base[-1]), K, 4))
    null_tab = np.zeros(int(base[-1]))
    for d in range(len(depths)):
        n = int(depths[d])
        cu, cf = np.zeros((n + 1, K, 4)), np.zeros((n // 2 + 1, K, 4))
        L.fo_sweep_model_tables(fsp[d].ctypes.data, n, K, asc_depth, asc_min_freq, ascbias_background_only,
                                include_invariant, cu.ctypes.data, cf.ctypes.data)
        coef[base[d]:base[d] + n + 1] = cu
        coef[base[d] + n + 1:base[d + 1]] = cf
        nu, nf = np.zeros(n + 1), np.zeros(n // 2 + 1)
        L.fo_null_model(fsp[d].ctypes.data, n, nu.ctypes.data, nf.ctypes.data)
        null_tab[base[d]:base[d] + n + 1] = nu
        null_tab[base[d] + n + 1:base[d + 1]] = nf
    cls = np.where(fo != 0, base[dp] + depths[dp] + 1 + of, base[dp] + of).astype(np.int32)
    chr_ = snp["chr"]
    n_chr = len(snp["chr_names"])
    start = np.array([int(np.argmax(chr_ == c)) for c in range(n_chr)], dtype=np.int32)
    cnt = np.array([int((chr_ == c).sum()) for c in range(n_chr)], dtype=np.int32)
    return dict(chr=np.ascontiguousarray(chr_), pos=np.ascontiguousarray(snp["pos"]), cls=cls,
                chr_names=snp["chr_names"], chr_start=start, chr_nsnps=cnt,
                chr_startpos=snp["pos"][start].astype(np.int32),
                chr_bplen=snp["pos"][start + cnt - 1].astype(np.int32),
                coef=coef, null_tab=null_tab, n_knots=K, step=24.0 / (K + 1.0), depths=depths, cls_base=base)


def table_text(table):
    names = table["chr_names"]
    c, p, k, d, f = (table[x] for x in ("chr", "pos", "count", "depth", "folded"))
    return "".join(f"{names[c[i]]}\t{p[i]}\t{k[i]}\t{d[i]}\t{f[i]}\n" for i in range(len(p)))
