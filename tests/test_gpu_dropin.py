"""Drop-in proof on the GPU: (1) the reference's unmodified main() + host objects
linked against libfscl_dropin.so (fscl_b200/_host/fscl_gpu) reproduce the
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
EXE = os.path.join(ROOT, "fscl_b200", "_host", "fscl_gpu")


@pytest.mark.parametrize("name", sorted(INDEX))
def test_reference_main_on_gpu_dropin_matches_golden(built, name, tmp_path):
    if not os.path.exists(EXE):
        pytest.fail("fscl_b200/_host/fscl_gpu missing: run build() where /root/reference exists")
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


def _run_dropin(inp, out, args, **env):
    cmd = [EXE, "-f", inp, "-o", out, "--n-threads=1", *args]
    e = dict(os.environ, OMP_NUM_THREADS="1", FSCL_B200_RNG="philox", FSCL_B200_SEED="7")
    e.update(env)
    return subprocess.run(cmd, env=e, capture_output=True, text=True)


def test_dropin_checkpoint_and_resume(built, tmp_path):
    """A permutation test stopped after a checkpoint and continued with FSCL_B200_RESUME writes
    the records of the uninterrupted run."""
    name = "unfolded_n40_perm"
    inp = str(tmp_path / "in.snps")
    open(inp, "w").write(gzip.open(os.path.join(GOLD, name + ".snps.gz"), "rt").read())
    args = ["-G", "100000", "-p", "300"]
    full, part, ck = str(tmp_path / "full.txt"), str(tmp_path / "part.txt"), str(tmp_path / "ck.bin")
    r = _run_dropin(inp, full, args)
    assert r.returncode == 0, r.stderr[-2000:]
    r = _run_dropin(inp, part, args, FSCL_B200_CHECKPOINT=ck, FSCL_B200_STOP_AFTER_ROUNDS="2")
    assert r.returncode == 3 and os.path.getsize(ck) > 0, r.stderr[-2000:]
    r = _run_dropin(inp, part, args, FSCL_B200_RESUME=ck)
    assert r.returncode == 0, r.stderr[-2000:]
    assert open(part).read() == open(full).read()
    # a checkpoint of another scan is refused
    r = _run_dropin(inp, part, [*args[:-1], "301"], FSCL_B200_RESUME=ck, FSCL_B200_SEED="8")
    assert r.returncode == 1 and "does not belong" in r.stderr


def test_dropin_sigint_writes_current_output(built, tmp_path):
    """SIGINT during the permutation test: current records + <out>-nulldist are written and the run
    goes on (scan-chromosome.c:557-569); a second signal inside the grace period aborts."""
    import signal
    import time
    name = "unfolded_n40_perm"
    inp, out = str(tmp_path / "in.snps"), str(tmp_path / "out.txt")
    open(inp, "w").write(gzip.open(os.path.join(GOLD, name + ".snps.gz"), "rt").read())
    args = ["-G", "100000", "-p", "100000"]
    # rounds of 4 permutations with a pause in between keep the test alive long enough to be interrupted
    env = dict(os.environ, OMP_NUM_THREADS="1", FSCL_B200_RNG="philox", FSCL_B200_SIGINT_GRACE="0.4",
               FSCL_B200_ROUND_PERMS="4", FSCL_B200_ROUND_DELAY_MS="100")
    p = subprocess.Popen([EXE, "-f", inp, "-o", out, "--n-threads=1", *args], env=env, stdout=subprocess.PIPE,
                         stderr=subprocess.PIPE, text=True)
    try:
        time.sleep(6.0)                                   # tables + forward scan + some rounds
        assert p.poll() is None, p.stderr.read()[-2000:]
        p.send_signal(signal.SIGINT)
        deadline = time.time() + 30
        while time.time() < deadline and not os.path.exists(out + "-nulldist"):
            time.sleep(0.1)
        assert os.path.exists(out + "-nulldist") and p.poll() is None
        time.sleep(1.0)
        rows = open(out).read().splitlines()
        assert rows and all(len(r.split("\t")) == 7 for r in rows)
        head = open(out + "-nulldist").readline().split("\t")
        assert head[:6] == ["chr", "pos", "CLR", "alpha", "p", "n"] and len(head) == 6 + 10000
        p.send_signal(signal.SIGINT)                      # two signals inside the grace period: abort
        time.sleep(0.05)
        p.send_signal(signal.SIGINT)
        assert p.wait(timeout=30) == 255
    finally:
        if p.poll() is None:
            p.kill()
