"""Regenerate tests/golden/: run the UNMODIFIED reference executable
(oracle/_ref/fscl_ref, built from /root/reference by oracle/Makefile) on small
synthetic SNP tables and store input + output.  Run in the build container only
(the GPU box has no /root/reference)."""
import gzip
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from fscl_b200 import synth  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def cases():
    rng = np.random.default_rng(17)
    t1 = synth.single_chromosome(2500, 1_500_000, 40, 21, planted=True)
    yield "unfolded_n40", t1, ["-G", "100000"], {}
    yield "unfolded_n40_perm", t1, ["-G", "100000", "-p", "30"], {}
    t2, _ = synth.config3(seed=9, scale=0.03, depths=(40, 38, 36, 34, 30))
    yield "asc1_folded_depths", t2, ["-G", "100000", "-d", "20", "--asc-minimum-freq=1"], dict(asc_depth=20, asc_min_freq=1)
    yield "asc1_folded_depths_perm", t2, ["-G", "100000", "-d", "20", "--asc-minimum-freq=1", "-p", "25"], \
        dict(asc_depth=20, asc_min_freq=1)
    pos = np.unique(np.concatenate([np.arange(1000, 1000 + 1500 * 40, 40), np.arange(1000, 61000, 280) + 1]))
    cnt = synth._neutral_counts(rng, len(pos), np.full(len(pos), 24))
    t3 = synth.make_table(["cA"], np.zeros(len(pos), dtype=np.int32), pos, cnt, np.full(len(pos), 24),
                          np.zeros(len(pos), dtype=np.int32))
    yield "coincident_grid", t3, ["-G", "4000", "--max-only"], {}
    yield "coincident_grid_all", t3, ["-G", "4000", "--prepend-label=lab"], {}


def main():
    os.makedirs(OUT, exist_ok=True)
    index = {}
    for name, tab, args, prep in cases():
        text = po.table_text(tab)
        with tempfile.TemporaryDirectory() as td:
            inp, out = os.path.join(td, "in.snps"), os.path.join(td, "out.txt")
            open(inp, "w").write(text)
            cmd = [po.REF_EXE, "-f", inp, "-o", out, "--n-threads=1", *args]
            subprocess.run(cmd, check=True, env=dict(os.environ, OMP_NUM_THREADS="1"),
                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            result = open(out).read()
        with gzip.open(os.path.join(OUT, name + ".snps.gz"), "wt") as fh:
            fh.write(text)
        open(os.path.join(OUT, name + ".out"), "w").write(result)
        index[name] = dict(args=args, prep=prep, n_records=result.count("\n"))
        print(name, "records", result.count("\n"))
    json.dump(index, open(os.path.join(OUT, "index.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
