#!/usr/bin/env python
"""bench.py -- permuted genome scans / s of the CLR sweep scan on B200.

Workload (BASELINE.json configs[1]): one 100 Mb chromosome, 200k SNPs, n=100,
-G 100000 (1000 cells, 12 CLR grid-evals each).  A *step* is one permutation
round: every rank builds R block-permuted genomes on its GPU and searches all
1000 cells of each (pruning off, so every scan is equal work); at N>1 the CLR
matrix of the round is summed over ranks with an NCCL allreduce (the real
sharded path of SURVEY.md 8e).  value = permuted scans / s over all ranks.

  python bench.py [--gpus N] [--steps K] [--warmup W]
  python bench.py --impl reference ...     # the unmodified reference on the host cores
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

G = 100_000
FLOP_PER_TERM = 12
WORKLOAD = "c2: 1 chr x 100 Mb, 200k SNPs, n=100, unfolded, -G 100000 (1000 cells x 12 grid-evals)"


def build_problem():
    from fscl_b200 import synth
    from oracle import pyoracle as po          # input preparation only (reference host code)
    tab, _ = synth.config2()
    rp = po.RefProblem(tab)
    return tab, rp, rp.flatten()


# ------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
            except (ValueError, IndexError):
                continue
            for nme, v in zip(names, r[2:6]):
                if v.strip().lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------- reference arm
def reference_cells_per_sec(rp, n_cells_target, threads, cells):
    """Time the reference's own search_maxpos (scan-chromosome.c:126) on
    evenly spread cells of the forward scan, `threads` host threads."""
    from oracle import pyoracle as po
    L = po.ref()
    s = rp.scan.contents
    idx = np.linspace(0, cells.shape[1] - 1, n_cells_target).astype(int)
    work = [(int(cells[0, i]), int(cells[1, i]), int(cells[2, i])) for i in idx]
    lock = threading.Lock()
    it = iter(work)

    def run():
        out = po.scan_pt_t()
        while True:
            with lock:
                w = next(it, None)
            if w is None:
                return
            L.fx_search_maxpos(C.byref(out), w[0], w[1], w[2], s.snps, C.byref(s.chr_limits[w[0]]),
                               81920, 128, rp.sm_p)

    ths = [threading.Thread(target=run) for _ in range(threads)]
    t0 = time.perf_counter()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    dt = time.perf_counter() - t0
    return len(work) / dt, dt


def enum_cells_host(flat):
    from oracle import pyoracle as po
    return po.Oracle(flat).enum_cells(G)


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    tab, rp, flat = build_problem()
    cells = enum_cells_host(flat)
    ncpu = len(os.sched_getaffinity(0))
    # a quarter of the scan per step: enough cells per thread that the tail of the last ones does not
    # understate the reference (16 cells on 16 threads measured 0.22 scans/s, 1000 cells 0.43)
    per_step = min(cells.shape[1], max(16 * ncpu, 256))
    for _ in range(args.warmup):
        reference_cells_per_sec(rp, max(ncpu // 2, 2), ncpu, cells)
    t_tot, c_tot = 0.0, 0
    for _ in range(args.steps):
        cps, dt = reference_cells_per_sec(rp, per_step, ncpu, cells)
        t_tot += dt
        c_tot += per_step
    scans = c_tot / cells.shape[1]
    v = scans / t_tot
    sample = f"{per_step} of {cells.shape[1]} cells of the forward scan per step, search_maxpos on {ncpu} threads"
    print(json.dumps({
        "impl": "reference", "metric": "permuted_genome_scans_per_sec", "value": v, "unit": "scans/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "grid_evals_per_sec": v * cells.shape[1] * 12,
        "cpu_baseline": {"value": v, "unit": "scans/s", "cores": ncpu, "kind": "reference", "sample": sample},
        "e2e": {"value": v, "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# ----------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="fscl_b200")
    ap.add_argument("--round-perms", type=int, default=16, help="permuted genomes per rank per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pruned-run", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import torch.distributed as dist
    from fscl_b200 import api

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (fscl_b200 has no CPU path)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    tab, rp, flat = build_problem()
    sc = api.Scanner(flat, device=local)
    obs = sc.scan_chromosome(G)
    n_cells = len(obs)
    R = args.round_perms
    matrix = torch.zeros(R * world * n_cells, dtype=torch.float64, device="cuda")

    def allreduce(m, n):
        if world > 1:
            dist.all_reduce(m[:n])
            torch.cuda.current_stream().synchronize()

    def run_rounds(k):
        """k steps = k rounds of R*world permutations, pruning off."""
        pts, st, _ = sc.scan_permute_sharded(obs, k * R * world - 1, G, rank, world, allreduce, matrix,
                                             matrix.data_ptr(), rng=api.RNG_PHILOX, seed=1, prune=0,
                                             round_perms=R * world)
        return pts, st

    # SNP-terms per round (roofline numerator), counted by an instrumented twin of the kernel:
    # the terms the kernel evaluates, and the terms of the reference algorithm for the same
    # scans (every coarse candidate summed site by site: SURVEY.md 8d's count)
    terms_round, _ = sc.bench_scan(G, R, count_terms=True)
    sc.set_exact_candidates(True)
    terms_round_ref, _ = sc.bench_scan(G, R, count_terms=True)
    sc.set_exact_candidates(False)
    fp64_peak = api.fp64_peak_tflops(local)

    clocks = ClockSampler(local)          # started before the warm-up: nvidia-smi needs ~0.2 s to deliver its first sample
    run_rounds(max(args.warmup, 3))
    sc.reset_stats()
    barrier()
    t0 = time.perf_counter()
    pts, st = run_rounds(args.steps)
    barrier()
    dt = time.perf_counter() - t0
    clk = clocks.stop()
    assert int(pts["permute_n"].min()) == args.steps * R * world
    tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt = float(tt.item())
    scans = args.steps * R * world
    value = scans / dt
    kern_ms = st["score_kernel_ms"] / max(st["score_launches"], 1)
    achieved_tf = FLOP_PER_TERM * terms_round / (kern_ms * 1e-3) / 1e12
    achieved_ref_tf = FLOP_PER_TERM * terms_round_ref / (kern_ms * 1e-3) / 1e12
    bound_failures = sc.stats()["bound_failures"]

    # ---- e2e: host buffers through the public API, H2D/D2H inside the timed region
    e2e_perms = 128         # permutations per end-to-end step (one H2D of the problem, one forward scan, results to host)

    e2e_phase = {"create_h2d": 0.0, "forward_scan": 0.0, "permuted_scans": 0.0, "close": 0.0}
    # the step's inputs live in pinned host memory (page-locked copies made once, outside the timed region)
    pinned = {k: torch.from_numpy(np.ascontiguousarray(flat[k])).pin_memory()
              for k in ("chr", "pos", "cls", "chr_start", "chr_nsnps", "chr_startpos", "chr_bplen", "coef", "null_tab")}
    flat_pinned = dict(flat)
    flat_pinned.update({k: v.numpy() for k, v in pinned.items()})

    def e2e_step():
        ta = time.perf_counter()
        s2 = api.Scanner(flat_pinned, device=local)
        tb = time.perf_counter()
        o = s2.scan_chromosome(G)
        tc = time.perf_counter()
        p, _, _ = s2.scan_permute(o, e2e_perms - 1, G, rng=api.RNG_PHILOX, seed=2, prune=0)
        td = time.perf_counter()
        s2.close()
        te_ = time.perf_counter()
        for k, v in zip(e2e_phase, (tb - ta, tc - tb, td - tc, te_ - td)):
            e2e_phase[k] += v
        return p
    # bytes that cross PCIe per step, counted from the buffers the library copies:
    #   in : the problem (sites, classes, spline table, null table), the log table, the cell task list,
    #        the scan points handed to the permutation test (chr, sweep_pos, clr, counters)
    #   out: 56-byte result per cell of the forward scan, 3 counters per cell, n_active + flags per round
    h2d = sum(flat[k].nbytes for k in ("chr", "pos", "cls", "coef", "null_tab")) + 65536 * 8 + n_cells * 20 + \
        n_cells * (5 * 4 + 8)
    d2h = n_cells * 56 + n_cells * 12 + 64
    e2e_step()
    barrier()
    for k in e2e_phase:
        e2e_phase[k] = 0.0
    t0 = time.perf_counter()
    n_e2e = 3
    for _ in range(n_e2e):
        e2e_step()
    barrier()
    e2e_dt = time.perf_counter() - t0
    te = torch.tensor([e2e_dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = n_e2e * (1 + e2e_perms) * world / float(te.item())

    extra = {}
    if rank == 0 and not args.no_pruned_run:
        # the configuration as quoted: 1000 permutations with stochastic 20-exceedance pruning (twice: the
        # first call also grows the permutation slots to the size a pruned test reserves)
        runs = []
        for _ in range(2):
            sc.reset_stats()
            t0 = time.perf_counter()
            p2, st2, _ = sc.scan_permute(obs, 1000, G, rng=api.RNG_PHILOX, seed=3, prune=1)
            runs.append(time.perf_counter() - t0)
        extra["pruned_run"] = {"n_permute": 1000, "seconds": min(runs), "first_call_seconds": runs[0],
                               "scan_kernel_ms": st2["score_kernel_ms"], "other_kernels_ms": st2["permute_kernel_ms"],
                               "rounds": st2["n_rounds"],
                               "scan_equivalents": st2["n_tasks"] / n_cells,
                               "cells_finished": int(p2["permute_finished"].sum()), "cells": n_cells}
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        cells = enum_cells_host(flat)
        ncpu = len(os.sched_getaffinity(0))
        cps0, _ = reference_cells_per_sec(rp, max(ncpu, 8), ncpu, cells)          # calibration + warm-up
        k = int(min(cells.shape[1], max(2 * ncpu, cps0 * 15.0)))                  # about 15 s of wall time
        cps, cdt = reference_cells_per_sec(rp, k, ncpu, cells)
        cpu = {"value": cps / cells.shape[1], "unit": "scans/s", "cores": ncpu, "kind": "reference",
               "sample": f"{k} of {cells.shape[1]} cells of the forward scan (a permuted scan is the same work), "
                         f"reference search_maxpos on {ncpu} threads, {cdt:.1f} s wall"}
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        traffic = None
        try:      # DRAM bytes of this kernel from the committed ncu capture, scaled to this launch's task count
            prof = json.load(open(os.path.join(ROOT, "profiles", "r01_scan_kernel.json")))
            traffic = prof["dram_bytes_per_task"] * R * n_cells
        except (OSError, KeyError, ValueError):
            pass
        line = {
            "metric": "permuted_genome_scans_per_sec", "value": value, "unit": "scans/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "permuted_genomes_per_rank_per_step": R, "pruning_in_timed_steps": False,
                       "rng": "philox4x32-10", "l2": "each step builds R new permuted genomes; the per-step working "
                       "set (R x 1.6 MB classes + 1 MB tables) is L2-resident by design (SURVEY.md 8d)"},
            "grid_evals_per_sec": value * n_cells * 12,
            "e2e": {"value": e2e_value, "unit": "scans/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "what": f"Scanner(host arrays) + scan_chromosome + scan_permute({e2e_perms} permutations) + results to host",
                    "phase_ms_per_step": {k: 1e3 * v / n_e2e for k, v in e2e_phase.items()}},
            "gpu_launches": int(st["kernel_launches"]),
            "clocks": clk,
            "roofline": {"bound": "fp64", "achieved": achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": achieved_tf / fp64_peak, "traffic": traffic,
                         "kernel": "scan_cells_kernel", "kernel_ms": kern_ms, "snp_terms_per_launch": int(terms_round),
                         "flop_per_term": FLOP_PER_TERM,
                         "what": "achieved/frac count the SNP-terms the kernel EVALUATES; coarse candidates whose sum "
                                 "is provably below the best exact one are excluded by an upper bound, not summed "
                                 "(results identical, DESIGN.md 3.1)",
                         "reference_equivalent": {"snp_terms_per_launch": int(terms_round_ref), "achieved": achieved_ref_tf,
                                                  "frac": achieved_ref_tf / fp64_peak,
                                                  "what": "the same launches credited with every term of the reference "
                                                          "algorithm (SURVEY.md 8d's count)"},
                         "coarse_passes_repeated_exactly": int(bound_failures),
                         "grid_evals_per_launch": int(R * n_cells * 12),
                         "peak_source": "FP64 FMA-chain microbenchmark run in this process (MEASURED_PEAKS.json "
                                        "has no FP64 figure; SURVEY.md 8d)",
                         "hbm_peak_gbs": peaks.get("hbm_gbs"),
                         "hbm_algorithmic_bytes_per_launch": int(R * (flat["pos"].nbytes * 2) + flat["coef"].nbytes + 524288),
                         "device_ms_per_step": (st["score_kernel_ms"] + st["permute_kernel_ms"]) / args.steps},
            "cpu_baseline": cpu,
        }
        line.update(extra)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
