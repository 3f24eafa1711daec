#!/usr/bin/env python
"""bench.py -- permuted genome scans / s of the CLR sweep scan + block-permutation test on B200.

Workload (BASELINE.json configs[3], the one north_star's target is quoted on; it fits one GPU): 22 autosomes,
3 M SNPs, n = 100, -G 100000 -> 28 790 cells x 12 CLR grid-evals per permuted genome scan.

* value      permuted scans / s, inputs resident in HBM, pruning OFF so that every scan is equal work:
             a *step* is one round of R block-permuted genomes per rank (built on the device: Philox block
             draws, swaps, class bucketing) with every cell of each searched; at N > 1 the permutations of
             a round shard over the ranks and one ncclAllGather per round (inside the library) completes the
             CLR matrix on every rank -- the real sharded path of SURVEY.md 8e.  weak scaling.
* e2e        the same metric through the public API from HOST buffers: Scanner(host arrays) [H2D of the
             problem] + forward scan + a sharded permutation test + counters back on the host, per step.
* pruned_run the configuration as north_star quotes it: 10 000 permutations WITH the stochastic
             20-exceedance pruning (scan-chromosome.c:488-498), strong-scaled over the N ranks through the
             same sharded path (fixed total work; wall seconds fall with N).  At N > 1 rank 0 repeats the
             test unsharded and asserts that permute_p / permute_n / permute_finished are identical.

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
WORKLOAD = "c4: 22 autosomes, 3M SNPs, n=100, unfolded, -G 100000 (28790 cells x 12 grid-evals per permuted scan)"
SEED_STEPS, SEED_E2E, SEED_PRUNED = 1, 2, 3


def config_dict(R):
    """identical in both arms (the driver compares the two lines' config)"""
    return {"workload": WORKLOAD, "permuted_genomes_per_rank_per_step": R, "pruning_in_timed_steps": False,
            "rng": "philox4x32-10",
            "l2": "inputs larger than L2 per step: each step builds R new permuted genomes of 60 MB each "
                  "(tiles + run lists + classes; R x 60 MB > 126 MB L2 at the default R = 4), nothing is reused "
                  "across steps except the 1 MB spline table and the 0.5 MB log table"}


def build_table():
    from fscl_b200 import synth
    tab, _ = synth.config4()
    return tab


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
class ReferenceCpu:
    """The reference's own code on the host cores (oracle/_ref = the unmodified reference compiled by
    oracle/Makefile): one block-permuted genome made by its snp_block_permute (scan-chromosome.c:336), then its
    search_maxpos (scan-chromosome.c:126) over a bounded sample of that genome's cells -- exactly what
    scan_permute_thread does per active cell (scan-chromosome.c:481-486) -- on all host threads."""

    def __init__(self, tab):
        from oracle import pyoracle as po
        self.po = po
        self.rp = po.RefProblem(tab)
        self.L = po.ref()
        flat = self.rp.flatten()
        self.cells = po.Oracle(flat).enum_cells(G)
        s = self.rp.scan.contents
        self.n_snps = s.n_snps
        self.p_snps = (po.snp_t * s.n_snps)()
        po.libc().srand(0xFD821A6)
        self.L.fx_snp_block_permute(self.p_snps, s.snps, s.n_snps, 0.1, 1.0)
        self.ncpu = len(os.sched_getaffinity(0))

    def cells_per_sec(self, n_cells):
        po, L, s = self.po, self.L, self.rp.scan.contents
        idx = np.linspace(0, self.cells.shape[1] - 1, n_cells).astype(int)
        work = []
        for i in idx:                                     # D4: the permuted cell is aligned to 0, not clipped
            chr_, start = int(self.cells[0, i]), int(self.cells[1, i])
            start -= start % G
            work.append((chr_, start, start + G))
        lock, it = threading.Lock(), iter(work)

        def run():
            out = po.scan_pt_t()
            while True:
                with lock:
                    w = next(it, None)
                if w is None:
                    return
                L.fx_search_maxpos(C.byref(out), w[0], w[1], w[2], self.p_snps, C.byref(s.chr_limits[w[0]]),
                                   81920, 128, self.rp.sm_p)

        ths = [threading.Thread(target=run) for _ in range(self.ncpu)]
        t0 = time.perf_counter()
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        dt = time.perf_counter() - t0
        return len(work) / dt, dt


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ref = ReferenceCpu(build_table())
    n_cells, ncpu = ref.cells.shape[1], ref.ncpu
    # enough cells per thread that the tail of the last ones does not understate the reference
    per_step = min(n_cells, max(16 * ncpu, 256))
    for _ in range(args.warmup):
        ref.cells_per_sec(max(ncpu // 2, 2))
    t_tot, c_tot = 0.0, 0
    for _ in range(args.steps):
        _, dt = ref.cells_per_sec(per_step)
        t_tot += dt
        c_tot += per_step
    v = (c_tot / n_cells) / t_tot
    sample = (f"{per_step} of the {n_cells} cells of one block-permuted genome per step (evenly spread over the 22 "
              f"chromosomes), reference snp_block_permute + search_maxpos on {ncpu} host threads")
    print(json.dumps({
        "impl": "reference", "metric": "permuted_genome_scans_per_sec", "value": v, "unit": "scans/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args.round_perms),
        "grid_evals_per_sec": v * n_cells * 12,
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
    ap.add_argument("--round-perms", type=int, default=4, help="permuted genomes per rank per step")
    ap.add_argument("--pruned-perms", type=int, default=10000, help="-p of the pruned, strong-scaled run")
    ap.add_argument("--e2e-perms", type=int, default=32, help="permutations per rank of one end-to-end step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pruned-run", action="store_true")
    ap.add_argument("--no-cli", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import torch.distributed as dist
    from fscl_b200 import api, hostref

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

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- the problem, prepared by the product's own host side (no oracle/): the reference's parser and
    # background spectrum (fscl_b200/host) + spline tables built on the GPU (fscl_tables.cu)
    t0 = time.perf_counter()
    tab = build_table()
    flat = hostref.load_problem(table=tab, device_tables=True, device=local)
    host_prep_s = time.perf_counter() - t0
    sc = api.Scanner(flat, device=local)
    if world > 1:
        # NCCL communicator INSIDE the library: rank 0 makes the id, torch.distributed only ships its 128 bytes
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid = torch.tensor(list(api.comm_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(uid, 0)
        sc.comm_init_rank(world, rank, bytes(uid.cpu().numpy().tobytes()))
    obs = sc.scan_chromosome(G)                      # sharded over cells at N > 1, identical on every rank
    n_cells = len(obs)
    R = args.round_perms

    def run_rounds(k):
        """k steps = k rounds of R * world permutations, pruning off (equal-work scans)."""
        pts, st, _ = sc.scan_permute(obs, k * R * world - 1, G, rng=api.RNG_PHILOX, seed=SEED_STEPS, prune=0,
                                     round_perms=R * world)
        return pts, st

    # SNP-terms per round (roofline numerator), counted by an instrumented twin of the kernel: the terms the
    # kernel evaluates, and -- for comparison with the CPU algorithm only -- the terms of the reference
    # algorithm for the same scans (every coarse candidate summed site by site: SURVEY.md 8d's count)
    terms_round, _ = sc.bench_scan(G, R, count_terms=True)
    sc.set_exact_candidates(True)
    terms_round_ref, _ = sc.bench_scan(G, 1, count_terms=True)
    terms_round_ref *= R
    sc.set_exact_candidates(False)
    fp64_peak = api.fp64_peak_tflops(local)

    clocks = ClockSampler(local)          # started before the warm-up: nvidia-smi needs ~0.2 s to deliver its first sample
    run_rounds(max(args.warmup, 3))
    sc.reset_stats()
    barrier()
    t0 = time.perf_counter()
    pts, st = run_rounds(args.steps)
    barrier()
    dt = max_over_ranks(time.perf_counter() - t0)
    clk = clocks.stop()
    assert int(pts["permute_n"].min()) == args.steps * R * world
    scans = args.steps * R * world
    value = scans / dt
    kern_ms = st["score_kernel_ms"] / max(st["score_launches"], 1)
    achieved_tf = FLOP_PER_TERM * terms_round / (kern_ms * 1e-3) / 1e12
    st_now = sc.stats()
    bound_failures, fine_failures = st_now["bound_failures"], st_now["fine_bound_failures"]
    near_ties = (st_now["near_ties_alpha"], st_now["near_ties_position"])
    launches_steps = int(st["kernel_launches"])

    # ---- e2e: host buffers through the public API, H2D/D2H inside the timed region.  One step per rank:
    # Scanner(pinned host arrays) + forward scan (cells sharded) + permutation test of e2e_perms * world
    # permutations (sharded) + counters back.  The communicator is handed from Scanner to Scanner.
    e2e_perms = args.e2e_perms
    e2e_phase = {"create_h2d": 0.0, "forward_scan": 0.0, "permuted_scans": 0.0, "close": 0.0}
    pinned = {k: torch.from_numpy(np.ascontiguousarray(flat[k])).pin_memory()
              for k in ("chr", "pos", "cls", "chr_start", "chr_nsnps", "chr_startpos", "chr_bplen", "coef", "null_tab")}
    flat_pinned = dict(flat)
    flat_pinned.update({k: v.numpy() for k, v in pinned.items()})
    holder = [sc]

    def e2e_step():
        ta = time.perf_counter()
        s2 = api.Scanner(flat_pinned, device=local)
        if world > 1:
            s2.comm_take(holder[0])
        tb = time.perf_counter()
        o = s2.scan_chromosome(G)
        tc = time.perf_counter()
        p, _, _ = s2.scan_permute(o, e2e_perms * world - 1, G, rng=api.RNG_PHILOX, seed=SEED_E2E, prune=0)
        td = time.perf_counter()
        if holder[0] is not sc:
            holder[0].close()
        holder[0] = s2
        te_ = time.perf_counter()
        for k, v in zip(e2e_phase, (tb - ta, tc - tb, td - tc, te_ - td)):
            e2e_phase[k] += v
        return p
    # bytes that cross PCIe per step and rank, counted from the buffers the library copies:
    #   in : the problem (sites, classes, spline table, null table), the log table, the cell task list,
    #        the scan points handed to the permutation test (chr, sweep_pos, clr, counters)
    #   out: 48-byte result per cell of the forward scan, 3 counters per cell, active counts + flags per round
    h2d = sum(flat[k].nbytes for k in ("chr", "pos", "cls", "coef", "null_tab")) + 65536 * 8 + n_cells * 20 + \
        n_cells * (5 * 4 + 8)
    d2h = n_cells * 48 + n_cells * 12 + 64
    e2e_step()
    barrier()
    for k in e2e_phase:
        e2e_phase[k] = 0.0
    t0 = time.perf_counter()
    n_e2e = 3
    for _ in range(n_e2e):
        e2e_step()
    barrier()
    e2e_dt = max_over_ranks(time.perf_counter() - t0)
    e2e_value = n_e2e * (1 + e2e_perms * world) / e2e_dt
    if world > 1:
        sc.comm_take(holder[0])
    if holder[0] is not sc:
        holder[0].close()

    # ---- the configuration as quoted: -p 10000 with stochastic pruning, strong-scaled over the ranks
    extra = {}
    if not args.no_pruned_run:
        # untimed: a short test that grows both slot sets of every rank to their full size (64 permuted genomes per rank,
        # 7.7 GB from the stream-ordered pool: a once-per-process cost of about a second)
        sc.scan_permute(obs, 64 * world - 1, G, rng=api.RNG_PHILOX, seed=SEED_PRUNED, prune=1)
        sc.reset_stats()
        barrier()
        t0 = time.perf_counter()
        p2, st2, _ = sc.scan_permute(obs, args.pruned_perms, G, rng=api.RNG_PHILOX, seed=SEED_PRUNED, prune=1)
        t_pr = max_over_ranks(time.perf_counter() - t0)
        barrier()
        tt = torch.tensor([float(st2["n_tasks"]), st2["score_kernel_ms"]], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt)
        needed = float(p2["permute_n"].sum())                 # searches the sequential rule needs: one per (cell, scored permutation)
        extra["pruned_run"] = {
            "n_permute": args.pruned_perms, "seconds": t_pr, "n_gpus": world, "sharded": world > 1,
            "path": "fscl_scan_permute over the library's NCCL communicator" if world > 1 else "fscl_scan_permute",
            "scan_equivalents": float(tt[0].item()) / n_cells, "searches": int(tt[0].item()),
            "searches_needed": int(needed), "wasted_search_fraction": float(tt[0].item()) / needed - 1.0,
            "rounds": int(st2["n_rounds"]), "scan_kernel_ms_sum_over_ranks": float(tt[1].item()),
            "cells_finished": int(p2["permute_finished"].sum()), "cells": n_cells,
            "limiter": "the scan kernel (seconds ~ scan_kernel_ms / n_gpus: rounds follow each other without idle "
                       "time on the scan stream); late rounds (a few hundred active cells x up to 64 permutations per "
                       "rank) run at 420 instead of 520 searches/ms because building up to 64 whole permuted genomes "
                       "per round shares the SMs with the scan"}
        if world > 1 and rank == 0:
            # N-GPU counts must equal the 1-GPU counts: the same test, unsharded, on rank 0's GPU
            s1 = api.Scanner(flat, device=local)
            s1.scan_permute(obs, 63, G, rng=api.RNG_PHILOX, seed=SEED_PRUNED, prune=1)     # untimed: its slot sets at full size too
            t0 = time.perf_counter()
            p1, _, _ = s1.scan_permute(obs, args.pruned_perms, G, rng=api.RNG_PHILOX, seed=SEED_PRUNED, prune=1)
            t_1 = time.perf_counter() - t0
            s1.close()
            same = all(np.array_equal(p1[k], p2[k]) for k in ("permute_p", "permute_n", "permute_finished"))
            assert same, "sharded permutation counts differ from the single-GPU counts"
            extra["pruned_run"]["equals_single_gpu_counts"] = bool(same)
            extra["pruned_run"]["single_gpu_seconds_same_box"] = t_1
            extra["pruned_run"]["strong_scaling_efficiency"] = t_1 / (world * t_pr)
        barrier()

    cpu = None
    cli = None
    if rank == 0 and world == 1 and not args.no_cli:
        # wall time of the product executable (the reference's own main() on the GPU drop-in) on this workload
        exe = hostref.FSCL_GPU_EXE
        if os.path.exists(exe):
            with tempfile.TemporaryDirectory() as td:
                f_in, f_out = os.path.join(td, "c4.snps"), os.path.join(td, "c4.out")
                hostref.write_snp_table(tab, f_in)
                env = dict(os.environ, FSCL_B200_RNG="philox", FSCL_B200_DEVICE_TABLES="1", OMP_NUM_THREADS="1")
                cmd = [exe, "-f", f_in, "-o", f_out, "-G", str(G), "-p", "1000", "-v", "0"]
                t0 = time.perf_counter()
                rcode = subprocess.run(cmd, env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL).returncode
                cli = {"seconds": time.perf_counter() - t0, "returncode": rcode,
                       "command": "fscl_gpu -f c4.snps -o out -G 100000 -p 1000 (FSCL_B200_RNG=philox, device tables)",
                       "records": sum(1 for _ in open(f_out)) if os.path.exists(f_out) else 0}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ref = ReferenceCpu(tab)
        cps0, _ = ref.cells_per_sec(max(ref.ncpu, 8))                              # calibration + warm-up
        k = int(min(ref.cells.shape[1], max(2 * ref.ncpu, cps0 * 15.0)))           # about 15 s of wall time
        cps, cdt = ref.cells_per_sec(k)
        cpu = {"value": cps / ref.cells.shape[1], "unit": "scans/s", "cores": ref.ncpu, "kind": "reference",
               "sample": f"{k} of the {ref.cells.shape[1]} cells of one block-permuted genome, reference snp_block_permute + "
                         f"search_maxpos on {ref.ncpu} threads, {cdt:.1f} s wall"}
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        traffic = None
        try:      # DRAM bytes of this kernel from the committed ncu --set full capture, per task, scaled to this launch
            prof = json.load(open(os.path.join(ROOT, "profiles", "r02_scan_kernel.json")))
            traffic = prof["dram_bytes_per_task"] * R * n_cells
        except (OSError, KeyError, ValueError):
            pass
        line = {
            "metric": "permuted_genome_scans_per_sec", "value": value, "unit": "scans/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(R),
            "grid_evals_per_sec": value * n_cells * 12,
            "grid_evals_note": "12 per cell as search_maxpos counts them (scan-chromosome.c:103-139); a cell that starts "
                               "where the previous one ended takes that end point's evaluation over, so about 11.1 per "
                               "cell are computed (bit-identical, fscl_set_endpoint_sharing)",
            "e2e": {"value": e2e_value, "unit": "scans/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "what": f"per rank and step: Scanner(pinned host arrays) + scan_chromosome (cells sharded) + scan_permute("
                            f"{e2e_perms} x {world} permutations, sharded, no pruning) + counters to host",
                    "phase_ms_per_step": {k: 1e3 * v / n_e2e for k, v in e2e_phase.items()}},
            "gpu_launches": launches_steps,
            "clocks": clk,
            "roofline": {"bound": "fp64", "achieved": achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": achieved_tf / fp64_peak, "traffic": traffic,
                         "kernel": "scan_cells_kernel", "kernel_ms": kern_ms, "snp_terms_per_launch": int(terms_round),
                         "flop_per_term": FLOP_PER_TERM,
                         "what": "achieved/frac count only the SNP-terms the kernel EVALUATES; coarse and fine candidates "
                                 "whose sum is provably below the best exact one are excluded by an upper bound, not summed "
                                 "(results identical, DESIGN.md 3.1). The excluded terms are the cheapest ones (far tiles, "
                                 "L1 hits), so every such exclusion lowers frac while it raises scans/s",
                         "coarse_passes_repeated_exactly": int(bound_failures),
                         "fine_passes_repeated_exactly": int(fine_failures),
                         "decisions_within_a_few_ulp_of_the_null_sum": {"alpha": int(near_ties[0]), "bisection": int(near_ties[1]),
                                                                       "of_grid_evals": int(args.steps * R * n_cells * 12)},
                         "grid_evals_per_launch": int(R * n_cells * 12),
                         "peak_source": "FP64 FMA-chain microbenchmark run in this process (MEASURED_PEAKS.json "
                                        "has no FP64 figure; SURVEY.md 8d)",
                         "hbm_peak_gbs": peaks.get("hbm_gbs"),
                         "hbm_algorithmic_bytes_per_launch": int(R * (flat["pos"].nbytes * 2) + flat["coef"].nbytes + 524288)},
            "algorithmic_speedup_not_a_roofline": {
                "reference_algorithm_snp_terms_per_launch": int(terms_round_ref),
                "evaluated_snp_terms_per_launch": int(terms_round),
                "terms_avoided_factor": terms_round_ref / max(terms_round, 1),
                "what": "how many SNP-terms the reference algorithm sums for the same scans vs how many this kernel "
                        "evaluates; explains the CPU ratio, says nothing about kernel efficiency"},
            "host_prep_s": host_prep_s,
            "cpu_baseline": cpu,
        }
        if cli:
            line["e2e_cli"] = cli
        line.update(extra)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
