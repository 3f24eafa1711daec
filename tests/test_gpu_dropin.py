"""Drop-in proof on the GPU: (1) the reference's unmodified main() + host objects
linked against libfscl_dropin.so (oracle/_ref/fscl_dropin_exe) reproduce the
golden outputs of the unmodified reference executable byte for byte, permutation
columns included (same libc rand() stream); (2) the Python host layer does the same."""
import gzip
import os
import subprocess

import numpy as np
import pytest

from tests.test_golden import GOLD, INDEX, load_case

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "fscl_dropin_exe")


@pytest.mark.parametrize("name", sorted(INDEX))
def test_reference_main_on_gpu_dropin_matches_golden(built, name, tmp_path):
    if not os.path.exists(EXE):
        pytest.fail("oracle/_ref/fscl_dropin_exe missing: run build() where /root/reference exists")
    inp, out = str(tmp_path / "in.snps"), str(tmp_path / "out.txt")
    open(inp, "w").write(gzip.open(os.path.join(GOLD, name + ".snps.gz"), "rt").read())
    cmd = [EXE, "-f", inp, "-o", out, "--n-threads=1", *INDEX[name]["args"]]
    r = subprocess.run(cmd, env=dict(os.environ, OMP_NUM_THREADS="1"), capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    assert open(out).read() == open(os.path.join(GOLD, name + ".out")).read()


@pytest.mark.parametrize("name", sorted(INDEX))
def test_python_host_layer_matches_golden(built, name):
    from fscl_b200 import api
    flat, opt, want = load_case(name)
    sc = api.Scanner(flat)
    pts = sc.scan_chromosome(opt["G"])
    if opt["p"]:
        pts, st, _ = sc.scan_permute(pts, opt["p"], opt["G"], rng=api.RNG_GLIBC, seed=0xFD821A6)
        assert st["d7_hits"] == 0
    got = sc.format_records(pts, opt["p"], opt["label"], maximum_only=opt["max_only"])
    assert got == want


def test_dropin_philox_mode_matches_api(built, tmp_path):
    """FSCL_B200_RNG=philox through the reference's main equals the API's Philox run."""
    from fscl_b200 import api
    name = "unfolded_n40_perm"
    flat, opt, _ = load_case(name)
    inp, out = str(tmp_path / "in.snps"), str(tmp_path / "out.txt")
    open(inp, "w").write(gzip.open(os.path.join(GOLD, name + ".snps.gz"), "rt").read())
    cmd = [EXE, "-f", inp, "-o", out, "--n-threads=1", *INDEX[name]["args"]]
    r = subprocess.run(cmd, env=dict(os.environ, OMP_NUM_THREADS="1", FSCL_B200_RNG="philox", FSCL_B200_SEED="99"),
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    sc = api.Scanner(flat)
    pts = sc.scan_chromosome(opt["G"])
    pts, _, _ = sc.scan_permute(pts, opt["p"], opt["G"], rng=api.RNG_PHILOX, seed=99)
    assert open(out).read() == sc.format_records(pts, opt["p"])
