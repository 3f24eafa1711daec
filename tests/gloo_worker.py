"""One rank of the 2-rank gloo test (tests/test_host_logic.py): the sharded
permutation protocol of fscl_b200.api.Scanner.scan_permute_sharded with the CLR
rows produced by the CPU oracle instead of the GPU."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from fscl_b200 import api  # noqa: E402
from oracle import pyoracle as po  # noqa: E402
from tests.test_golden import load_case  # noqa: E402
from tests.test_host_logic import _oracle_clr_matrix  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    flat, opt, _ = load_case("unfolded_n40_perm")
    G, P, seed, R = opt["G"], 30, 5, int(os.environ.get("FSCL_TEST_ROUND", "6"))
    o = po.Oracle(flat, mode=2)
    pts0, _ = o.scan(G)
    pts = np.zeros(len(pts0), dtype=api.PT_DTYPE)
    for k in ("chr", "sweep_pos", "clr", "lalpha"):
        pts[k] = pts0[k]
    gather = len(sys.argv) > 2 and sys.argv[2] == "gather"
    t = 0
    while t <= P:
        active = np.nonzero(pts["permute_finished"] == 0)[0]
        if len(active) == 0:
            break
        r = min(R, P + 1 - t)
        mine = [q for q in range(t, t + r) if q % world == rank]          # this rank's permutations
        if gather:
            # the library's protocol: every rank owns a chunk [ns_max][n_active] of the gathered matrix and fills its
            # own rows only; one all-gather completes it (fscl_host.cu: device_round_commit)
            ns_max = (r + world - 1) // world
            chunk = ns_max * len(active)
            own = np.zeros(chunk)
            if mine:
                rows = _oracle_clr_matrix(o, pts0, G, seed, mine)
                for q, row in enumerate(rows):
                    own[q * len(active):(q + 1) * len(active)] = row[active]
            parts = [torch.zeros(chunk, dtype=torch.float64) for _ in range(world)]
            dist.all_gather(parts, torch.from_numpy(own))
            api.host_replay_gathered(pts, active, torch.cat(parts).numpy(), world, chunk, r, t, seed=seed)
            t += r
            continue
        m = np.zeros((r, len(active)))
        if mine:
            rows = _oracle_clr_matrix(o, pts0, G, seed, mine)
            for q, row in zip(mine, rows):
                m[q - t] = row[active]
        mt = torch.from_numpy(m)
        dist.all_reduce(mt)                                                # exceedance / CLR matrix, summed over ranks
        api.host_replay_round(pts, active, mt.numpy(), t, seed=seed)
        t += r
    np.save(sys.argv[1], np.stack([pts["permute_p"], pts["permute_n"]]))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
