"""The product's host side (fscl_b200/host, fscl_b200/hostref.py) on the CPU: the loader gives the oracle's
problem bit for bit, and the C ms reader (ms_reader.c, behind the reference's ms_openfile / ms_background /
ms_next_block) gives the table the Python statement of the same semantics (msformat.py) gives."""
import ctypes as C
import os

import numpy as np
import pytest

from fscl_b200 import hostref, msformat, synth
from oracle import pyoracle as po

FLAT_KEYS = ("chr", "pos", "cls", "chr_start", "chr_nsnps", "chr_startpos", "chr_bplen", "coef", "null_tab")


@pytest.fixture(scope="module")
def host(built):
    if not os.path.exists(hostref.HOSTREF_SO):
        pytest.fail("fscl_b200/_host/libfscl_hostref.so missing: run build() where /root/reference exists")
    return hostref.load_library()


def test_loader_equals_reference_preparation(host):
    """load_problem(device_tables=False) = load_snp_input + background_fsp + compute_sweep_model_tables +
    compute_snp_null_model of the unmodified reference (oracle/_ref), array for array."""
    tab = synth.single_chromosome(4000, 2_000_000, 50, 11, planted=True)
    a, b = hostref.load_problem(table=tab, device_tables=False), po.RefProblem(tab).flatten()
    for k in FLAT_KEYS:
        assert np.array_equal(a[k], b[k]), k
    assert a["chr_names"] == b["chr_names"] and a["n_knots"] == b["n_knots"] and a["step"] == b["step"]
    tab, opt = synth.config3(scale=0.03)
    a = hostref.load_problem(table=tab, device_tables=False, asc_depth=20, asc_min_freq=2)
    b = po.RefProblem(tab, asc_depth=20, asc_min_freq=2).flatten()
    for k in FLAT_KEYS:
        assert np.array_equal(a[k], b[k], equal_nan=True), k


def test_loader_reads_a_file_and_rejects_bad_calls(host, tmp_path):
    tab = synth.single_chromosome(500, 200_000, 20, 3)
    path = tmp_path / "in.snps"
    hostref.write_snp_table(tab, str(path))
    a = hostref.load_problem(snp_file=str(path), device_tables=False)
    b = hostref.load_problem(table=tab, device_tables=False)
    assert all(np.array_equal(a[k], b[k]) for k in FLAT_KEYS)
    with pytest.raises(ValueError):
        hostref.load_problem()
    with pytest.raises(ValueError):
        hostref.load_problem(snp_file=str(path), table=tab)


def _ms_blocks(rng, n_blocks=4, n_hap=14, S=60):
    out = []
    for _ in range(n_blocks):
        pos = np.sort(np.round(rng.random(S), 6))
        hap = (rng.random((n_hap, S)) < rng.random(S)).astype(np.uint8)
        out.append(dict(positions=pos, haplotypes=hap))
    out.insert(2, dict(positions=np.zeros(0), haplotypes=np.zeros((0, 0), dtype=np.uint8)))     # segsites: 0
    return out


@pytest.mark.parametrize("folded,first,size", [(0, 0, 0), (1, 0, 0), (0, 2, 8), (1, 3, 10)])
def test_c_ms_reader_equals_python_statement(host, tmp_path, folded, first, size):
    rng = np.random.default_rng(21)
    blocks = _ms_blocks(rng)
    text = msformat.write_ms(blocks, 14)
    text = text.replace("//\n", "//\n[14](1:0.1,2:0.1);\n", 1)                # a tree line is skipped
    path = tmp_path / "in.ms"
    path.write_text(text)
    L = 250_000
    want = msformat.ms_to_table(msformat.parse_ms(text), L, folded=bool(folded), sample_first=first, sample_size=size)
    so = host.fscl_ms_read(str(path).encode(), L, folded, first, size)
    s = so.contents
    snps = np.frombuffer((C.c_char * (32 * s.n_snps)).from_address(C.addressof(s.snps.contents)), dtype=hostref.SNP_DTYPE)
    assert s.n_snps == len(want["pos"]) and s.n_chromosomes == len(want["chr_names"])
    assert [s.chr_limits[i].name.decode() for i in range(s.n_chromosomes)] == want["chr_names"]
    assert np.array_equal(snps["chr"], want["chr"]) and np.array_equal(snps["pos"], want["pos"])
    assert np.array_equal(snps["obs_freq"], want["count"]) and np.array_equal(snps["folded"], want["folded"])
    depths = [s.sample_depths[i] for i in range(s.n_depths)]
    assert np.array_equal(np.array(depths)[snps["depth_p"]], want["depth"])
    for i in range(s.n_chromosomes):
        lim = s.chr_limits[i]
        sel = want["chr"] == i
        assert lim.n_snps == int(sel.sum()) and lim.start_pos == want["pos"][sel][0] and lim.bp_length == want["pos"][sel][-1]
        assert lim.start_index == int(np.argmax(sel))
    host.fscl_ms_free(so)


def test_ms_entry_points_follow_the_reference_protocol(host, tmp_path):
    """fscl.c:281-310: ms_background (all blocks), then ms_openfile + ms_next_block until NULL.  All blocks come
    back as ONE scan object (one chromosome per block), so the loop body runs once."""
    rng = np.random.default_rng(22)
    path = tmp_path / "in.ms"
    path.write_text(msformat.write_ms(_ms_blocks(rng), 14))
    host.ms_background.restype = C.POINTER(hostref.scan_t)
    host.ms_background.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int]
    host.ms_next_block.restype = C.POINTER(hostref.scan_t)
    host.ms_next_block.argtypes = [C.c_int] * 4
    host.ms_openfile.argtypes = [C.c_char_p]
    bg = host.ms_background(str(path).encode(), 100_000, 0, 0, 0)
    host.ms_openfile(str(path).encode())
    first = host.ms_next_block(100_000, 0, 0, 0)
    assert bool(first) and first.contents.n_snps == bg.contents.n_snps and first.contents.n_chromosomes == 4
    assert not bool(host.ms_next_block(100_000, 0, 0, 0))
    host.ms_openfile(str(path).encode())                                        # re-opening starts over
    again = host.ms_next_block(100_000, 0, 0, 0)
    assert bool(again) and again.contents.n_snps == bg.contents.n_snps
    for so in (bg, first, again):
        host.fscl_ms_free(so)


@pytest.mark.parametrize("text,why", [
    ("//\nsegsites: 3\npositions: 0.1 0.2\n010\n", "positions do not match segsites"),
    ("//\nsegsites: 2\npositions: 0.1 0.2\n0101\n", "haplotype row does not match segsites"),
    ("//\nsegsites: 2\npositions: 0.1 0.2\n01\nhello\n", "unexpected text"),
    ("//\npositions: 0.1\n1\n", "positions before segsites"),
])
def test_c_ms_reader_rejects_malformed_input(host, tmp_path, text, why):
    """Errors follow the reference's convention (message on stderr, exit(1)), so the reader is driven in a child process."""
    import subprocess
    import sys
    path = tmp_path / "bad.ms"
    path.write_text(text)
    code = ("import sys; sys.path.insert(0, %r); from fscl_b200 import hostref; L = hostref.load_library(); "
            "L.fscl_ms_read(%r.encode(), 1000, 0, 0, 0)" % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), str(path)))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode == 1 and why in r.stderr, (r.returncode, r.stderr[-300:])
