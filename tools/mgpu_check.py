"""Multi-GPU parity check, run under torchrun (one process per GPU):
neighbourhood-sharded best-improvement 2-opt (tiles dealt round-robin over the ranks, the per-rank argmin keys exchanged
through NVLink peer slots by the scan kernel, or by one 8-byte NCCL min-allreduce per pass) — exhaustive and with exact tile
pruning — and the sharded first-improvement search must produce exactly the single-GPU move log, tour and cost on every
rank."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from tsp_optimization_b200 import BI, FI, Engine  # noqa: E402
from tsp_optimization_b200.dist import attach_engine_comm, init_process_group_from_env  # noqa: E402
from tsp_optimization_b200.instances import uniform_instance  # noqa: E402


def main():
    rank, world, local = init_process_group_from_env("nccl")
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 6000
    passes = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    wt = 0
    xy = uniform_instance(n)
    if len(sys.argv) > 3:  # a fixture instance by name, e.g. gr666 (GEO: the exact scan, keys min-allreduced by NCCL)
        z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "instances.npz"))
        xy, wt = z[sys.argv[3] + "__xy"], int(z[sys.argv[3] + "__wt"])
        n = len(xy)
    single = Engine(local)
    single.set_instance(xy, wt)
    succ0, c0 = single.nn_tour(0)
    single.set_option("prune", 0)
    single.set_option("single_block", 0)
    s1, o1, st1, log1 = single.two_opt(BI, succ0, 0.0, max_iters=passes, log_cap=passes + 8)
    fi_moves = 3 * passes
    f1, fo1, fst1, flog1 = single.two_opt(FI, succ0, c0, max_iters=fi_moves, log_cap=fi_moves + 8)
    single.close()

    eng = Engine(local)
    eng.set_instance(xy, wt)
    attach_engine_comm(eng, rank, world)
    ok = True
    times = {}
    eng.set_option("single_block", 0)
    for exchange, name in ((0, "p2p"), (1, "nccl")):
        eng.set_option("exchange", exchange)
        for prune in (0, 1):
            eng.set_option("prune", prune)
            s2, o2, st2, log2 = eng.two_opt(BI, succ0, 0.0, max_iters=passes, log_cap=passes + 8)
            times[name + ("_pruned" if prune else "")] = st2.gpu_ms
            ok = ok and (s1 == s2).all() and o1 == o2 and log1.tolist() == log2.tolist() and st1.passes == st2.passes
        # first improvement: the segments of the pair order dealt over the ranks — for every search (gap 0), only after a
        # search longer than 20 000 pairs (the decision flips many times in a run of this size), and with the default threshold
        for gap, tag in ((0, ""), (20000, "_mixed"), (4000000, "_auto")):
            eng.set_option("fi_shard_min_gap", gap)
            f2, fo2, fst2, flog2 = eng.two_opt(FI, succ0, c0, max_iters=fi_moves, log_cap=fi_moves + 8)
            times["fi_" + name + tag] = fst2.gpu_ms
            ok = ok and (f1 == f2).all() and fo1 == fo2 and flog1.tolist() == flog2.tolist() and fst1.moves == fst2.moves
    eng.set_option("exchange", 0)
    eng.set_option("prune", -1)
    p2p = eng.info("exchange_p2p")
    # all ranks must agree with each other as well
    h = torch.tensor([int(np.int64(np.sum(s2.astype(np.int64) * np.arange(1, n + 1))) % (1 << 62)), int(ok)], dtype=torch.int64, device="cuda")
    hs = [torch.zeros_like(h) for _ in range(world)]
    dist.all_gather(hs, h)
    same = all(int(x[0]) == int(hs[0][0]) for x in hs) and all(int(x[1]) == 1 for x in hs)
    if rank == 0:
        print(f"MGPU_CHECK world={world} n={n} passes={st2.passes} moves={st2.moves} "
              f"single_ms={st1.gpu_ms:.2f} sharded_p2p_ms={times['p2p']:.2f} sharded_nccl_ms={times['nccl']:.2f} "
              f"pruned_p2p_ms={times['p2p_pruned']:.2f} fi_moves={fst2.moves} fi_single_ms={fst1.gpu_ms:.2f} fi_p2p_ms={times['fi_p2p']:.2f} "
              f"fi_nccl_ms={times['fi_nccl']:.2f} fi_p2p_mixed_ms={times['fi_p2p_mixed']:.2f} fi_p2p_auto_ms={times['fi_p2p_auto']:.2f} "
              f"p2p_enabled={p2p} {'OK' if same else 'MISMATCH'}", flush=True)
    eng.close()
    dist.destroy_process_group()
    sys.exit(0 if same else 1)


if __name__ == "__main__":
    main()
