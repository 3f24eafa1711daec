"""c5 at full size through the product executable: 20 M SNPs, -G 10000 (287 900 cells), -p N with the counter-based stream,
<out>-nulldist and <out>-projected written; reports wall time and the peak resident set of the process (VERDICT r1 #7:
stored CLRs must not cost 40 KB per cell on the host)."""
import json, os, resource, subprocess, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fscl_b200 import hostref, synth
perms = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
t0 = time.time()
tab, opt = synth.config5(scale=scale)
with tempfile.TemporaryDirectory() as td:
    f_in, f_out = os.path.join(td, "c5.snps"), os.path.join(td, "c5.out")
    hostref.write_snp_table(tab, f_in)
    n_sites = len(tab["pos"])
    del tab
    t_gen = time.time() - t0
    env = dict(os.environ, FSCL_B200_RNG="philox", FSCL_B200_DEVICE_TABLES="1", FSCL_B200_NULLDIST="1", FSCL_B200_STATS="1", OMP_NUM_THREADS="1")
    t0 = time.time()
    r = subprocess.run([hostref.FSCL_GPU_EXE, "-f", f_in, "-o", f_out, "-G", str(opt["G"]), "-p", str(perms), "-v", "0"], env=env,
                       capture_output=True, text=True)
    wall = time.time() - t0
    rss_mb = resource.getrusage(resource.RUSAGE_CHILDREN).ru_maxrss / 1024.0
    out = {"workload": f"c5 x {scale}: {n_sites} SNPs, -G {opt['G']}, -p {perms}, FSCL_B200_RNG=philox, device tables, nulldist + projected written",
           "returncode": r.returncode, "wall_s": wall, "peak_rss_mb": rss_mb, "input_generation_s": t_gen,
           "records": sum(1 for _ in open(f_out)) if os.path.exists(f_out) else 0,
           "nulldist_mb": os.path.getsize(f_out + "-nulldist") / 1e6 if os.path.exists(f_out + "-nulldist") else None,
           "projected_lines": sum(1 for _ in open(f_out + "-projected")) if os.path.exists(f_out + "-projected") else None,
           "stderr_tail": r.stderr[-400:]}
    print(json.dumps(out))
