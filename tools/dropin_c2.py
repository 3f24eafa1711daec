"""The reference's own main() linked against the drop-in (fscl_b200/_host/fscl_gpu) on c2 with 1000 permutations:
wall time of the whole command line, counter-based stream and the reference's rand() stream."""
import os, subprocess, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fscl_b200 import synth
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
exe = os.path.join(ROOT, "fscl_b200", "_host", "fscl_gpu")
tab, opt = synth.config2()
d = tempfile.mkdtemp()
inp = os.path.join(d, "c2.snps")
synth.write_snp_table(tab, inp)
for name, env in (("philox", {"FSCL_B200_RNG": "philox"}), ("libc rand() stream (reference-identical counts)", {})):
    out = os.path.join(d, "out_" + name.split()[0])
    t0 = time.perf_counter()
    r = subprocess.run([exe, "-f", inp, "-o", out, "-G", "100000", "-p", "1000", "--n-threads=1"],
                       env=dict(os.environ, OMP_NUM_THREADS="1", **env), capture_output=True, text=True)
    dt = time.perf_counter() - t0
    rows = open(out).read().splitlines() if os.path.exists(out) else []
    print(f"{name}: rc {r.returncode}, {dt:.2f} s wall, {len(rows)} records; first: {rows[0] if rows else r.stderr[-300:]}")
