"""Input preparation for the GPU scan: SNP / ms files -> the flattened problem of the C-ABI.

The host side of the product (north_star: "C host code reuses the input parsing, background-fsp.c
and cmdline-utils.c"): ``fscl_b200/_host/libfscl_hostref.so`` is built by ``fscl_b200/host/Makefile``
from the reference's host files where they lie (``load_snp_input`` snp-input.c:19,
``background_fsp`` background-fsp.c:182) plus this repository's ms reader
(``fscl_b200/host/ms_reader.c``).  The sweep-model spline tables and the null model come from
``fscl_build_tables`` (CUDA, ``fscl_tables.cu``) by default, or from the reference's own
``compute_sweep_model_tables`` (sm-spline.c:486, bit-exact tables) with ``device_tables=False``.

Nothing here imports ``oracle/``.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import tempfile

import numpy as np

from . import api

HERE = os.path.dirname(os.path.abspath(__file__))
HOSTREF_SO = os.path.join(HERE, "_host", "libfscl_hostref.so")
FSCL_GPU_EXE = os.path.join(HERE, "_host", "fscl_gpu")


# struct mirrors of include/fscl_dropin.h (= fscl.h:7-76)
class snp_t(C.Structure):
    _fields_ = [("chr", C.c_int), ("pos", C.c_int), ("null_logl", C.c_double),
                ("obs_freq", C.c_int), ("depth_p", C.c_int), ("folded", C.c_int)]


SNP_DTYPE = np.dtype({"names": ["chr", "pos", "null_logl", "obs_freq", "depth_p", "folded"],
                      "formats": ["<i4", "<i4", "<f8", "<i4", "<i4", "<i4"],
                      "offsets": [0, 4, 8, 16, 20, 24], "itemsize": 32})


class chr_limits_t(C.Structure):
    _fields_ = [("chr", C.c_int), ("name", C.c_char_p), ("start_index", C.c_int),
                ("n_snps", C.c_int), ("start_pos", C.c_int), ("bp_length", C.c_int)]


class scan_t(C.Structure):
    _fields_ = [("n_snps", C.c_int), ("snps", C.POINTER(snp_t)), ("n_depths", C.c_int),
                ("sample_depths", C.POINTER(C.c_int)), ("n_scan_pts", C.c_int),
                ("scan_pts", C.c_void_p), ("chr_limits", C.POINTER(chr_limits_t)),
                ("n_chromosomes", C.c_int)]


class spline_t(C.Structure):
    _fields_ = [("n", C.c_int), ("knot_points", C.POINTER(C.c_double)),
                ("coef", C.POINTER(C.POINTER(C.c_double)))]


class sm_ptable_t(C.Structure):
    _fields_ = [("spline_func", C.POINTER(C.POINTER(spline_t))),
                ("fspline_func", C.POINTER(C.POINTER(spline_t))),
                ("sample_size", C.c_int), ("pbk", C.POINTER(C.POINTER(C.c_double))),
                ("fsp", C.POINTER(C.c_double))]


_lib = None


def _libm_log(x):
    """log() of the C library, as compute_snp_null_model calls it (scan-chromosome.c:31-34); numpy's vectorised
    log differs in the last ulp on some CPUs"""
    x = float(x)
    if x > 0.0:
        return math.log(x)
    return -math.inf if x == 0.0 else math.nan


def load_library(path=HOSTREF_SO):
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise api.FsclError(f"{path} not built: run `make -C fscl_b200/host` where the reference tree is "
                            "available (the prebuilt library travels with the repository snapshot)")
    L = C.CDLL(path)
    L.load_snp_input.restype = C.POINTER(scan_t)
    L.load_snp_input.argtypes = [C.c_char_p, C.c_int, C.c_int]
    L.background_fsp.restype = C.POINTER(C.POINTER(C.c_double))
    L.background_fsp.argtypes = [C.POINTER(scan_t), C.c_int, C.c_char_p, C.c_int]
    L.ref_compute_sweep_model_tables.restype = C.POINTER(sm_ptable_t)
    L.ref_compute_sweep_model_tables.argtypes = [C.POINTER(scan_t), C.POINTER(C.POINTER(C.c_double)),
                                                 C.c_int, C.c_int, C.c_int, C.c_int]
    L.fscl_ms_read.restype = C.POINTER(scan_t)
    L.fscl_ms_read.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int]
    L.fscl_ms_free.argtypes = [C.POINTER(scan_t)]
    L.configure_logmsg.argtypes = [C.c_int]
    L.configure_logmsg(1)
    os.environ.setdefault("OMP_NUM_THREADS", "1")     # the reference's OpenMP loops race
    _lib = L
    return L


class _quiet_stderr:
    """load_snp_input prints progress with plain fprintf(stderr) (snp-input.c:112-118)."""

    def __enter__(self):
        self.null = os.open(os.devnull, os.O_WRONLY)
        self.saved = os.dup(2)
        os.dup2(self.null, 2)

    def __exit__(self, *a):
        os.dup2(self.saved, 2)
        os.close(self.saved)
        os.close(self.null)


def write_snp_table(table, path):
    """5-column text the reference parses (snp-input.c:49-66)."""
    names = table["chr_names"]
    c, p, k, d, f = (np.asarray(table[x]) for x in ("chr", "pos", "count", "depth", "folded"))
    nm = np.asarray(names, dtype=object)[c]
    with open(path, "w") as fh:
        fh.write("\n".join(f"{a}\t{b}\t{cc}\t{dd}\t{e}" for a, b, cc, dd, e in
                           zip(nm, p.tolist(), k.tolist(), d.tolist(), f.tolist())))
        fh.write("\n")


def load_problem(snp_file=None, table=None, ms_file=None, ms_segment_length=0, ms_folded=0, ms_sample_first=0,
                 ms_sample_size=0, asc_depth=0, asc_min_freq=1, ascbias_background_only=0, include_invariant=0,
                 force_neutral=0, minimum_depth=5, spline_pts=200, device_tables=True, device=0, quiet=True):
    """fscl.c:316-326 for the GPU: parse, background spectrum, tables, null model -> ``flat`` dict
    (the arrays ``api.Scanner`` takes; see ``fscl_problem_t`` in include/fscl_b200.h).

    Exactly one of ``snp_file`` (5-column SNP table), ``table`` (dict of arrays, written to a
    temporary file and parsed by the same code) or ``ms_file`` (Hudson ms text) is given."""
    L = load_library()
    if sum(x is not None for x in (snp_file, table, ms_file)) != 1:
        raise ValueError("give exactly one of snp_file, table, ms_file")
    C.c_int.in_dll(L, "spline_pts").value = spline_pts
    tmp = None
    try:
        if table is not None:
            fd, tmp = tempfile.mkstemp(suffix=".snps")
            os.close(fd)
            write_snp_table(table, tmp)
            snp_file = tmp
        if ms_file is not None:
            scan = L.fscl_ms_read(os.fsencode(ms_file), ms_segment_length, ms_folded, ms_sample_first, ms_sample_size)
        elif quiet:
            with _quiet_stderr():
                scan = L.load_snp_input(os.fsencode(snp_file), include_invariant, minimum_depth)
        else:
            scan = L.load_snp_input(os.fsencode(snp_file), include_invariant, minimum_depth)
    finally:
        if tmp:
            os.unlink(tmp)
    s = scan.contents
    fsp_p = L.background_fsp(scan, force_neutral, None, include_invariant)
    depths = np.array([s.sample_depths[i] for i in range(s.n_depths)], dtype=np.int32)
    fsp = [np.array([fsp_p[d][f] for f in range(int(depths[d]) + 1)]) for d in range(s.n_depths)]
    base = np.zeros(s.n_depths + 1, dtype=np.int64)
    for d in range(s.n_depths):
        base[d + 1] = base[d] + (depths[d] + 1) + (depths[d] // 2 + 1)
    n_cls, K = int(base[-1]), spline_pts
    if device_tables:
        coef, null_tab, flags = api.build_tables(depths, fsp, spline_pts, asc_depth, asc_min_freq,
                                                 ascbias_background_only, include_invariant, device=device)
    else:
        sm_p = L.ref_compute_sweep_model_tables(scan, fsp_p, asc_depth, asc_min_freq, ascbias_background_only,
                                                include_invariant)
        coef = np.zeros((n_cls, K, 4), dtype=np.float64)
        null_tab = np.zeros(n_cls, dtype=np.float64)
        flags = 0
        for d in range(s.n_depths):
            n = int(depths[d])
            t = sm_p[d]
            for f in range(n + 1):
                coef[base[d] + f] = np.ctypeslib.as_array(t.spline_func[f].contents.coef[0], shape=(K, 4))
                null_tab[base[d] + f] = _libm_log(fsp[d][f])                                 # scan-chromosome.c:23-37
            for f in range(n // 2 + 1):
                coef[base[d] + n + 1 + f] = np.ctypeslib.as_array(t.fspline_func[f].contents.coef[0], shape=(K, 4))
                null_tab[base[d] + n + 1 + f] = (_libm_log(fsp[d][f] + fsp[d][n - f]) if f != n - f
                                                 else _libm_log(fsp[d][f]))
    buf = (C.c_char * (32 * s.n_snps)).from_address(C.addressof(s.snps.contents))
    snps = np.frombuffer(buf, dtype=SNP_DTYPE)
    cls = np.where(snps["folded"] != 0,
                   base[snps["depth_p"]] + depths[snps["depth_p"]] + 1 + snps["obs_freq"],
                   base[snps["depth_p"]] + snps["obs_freq"]).astype(np.int32)
    lim = [s.chr_limits[i] for i in range(s.n_chromosomes)]
    return dict(
        chr=np.array(snps["chr"]), pos=np.array(snps["pos"]), cls=cls,
        chr_names=[l.name.decode() for l in lim],
        chr_start=np.array([l.start_index for l in lim], dtype=np.int32),
        chr_nsnps=np.array([l.n_snps for l in lim], dtype=np.int32),
        chr_startpos=np.array([l.start_pos for l in lim], dtype=np.int32),
        chr_bplen=np.array([l.bp_length for l in lim], dtype=np.int32),
        coef=coef, null_tab=null_tab, n_knots=K, step=24.0 / (K + 1.0),
        depths=depths, cls_base=base, fsp=fsp, table_flags=flags,
        obs_freq=np.array(snps["obs_freq"]), depth_p=np.array(snps["depth_p"]), folded=np.array(snps["folded"]),
    )
