"""BASELINE configs[4] across GPUs: 1024 random 1000-node tours (a GA population) repaired by 2-opt, the batch split
contiguously over the ranks (tsp_optimization_b200.dist.shard_batch), no data-path collective.  Run under torchrun; every
rank checks its shard against a single-GPU run of the same tours, rank 0 prints one JSON line."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from tsp_optimization_b200 import BI, FI, Engine  # noqa: E402
from tsp_optimization_b200.dist import init_process_group_from_env, shard_batch  # noqa: E402
from tsp_optimization_b200.instances import random_tours, uniform_instance  # noqa: E402


def main():
    rank, world, local = init_process_group_from_env("nccl")
    n, B = 1000, 1024
    mode = BI if (len(sys.argv) > 1 and sys.argv[1] == "BI") else FI
    xy = uniform_instance(n)
    tours = random_tours(n, B, 7)
    eng = Engine(local)
    eng.set_instance(xy, 0)
    costs = eng.tour_costs(tours, as_order=False)
    lo, hi = shard_batch(B, rank, world)
    eng.two_opt_batch(mode, tours[lo:lo + 2], costs[lo:lo + 2])  # warm-up
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    s, o, st = eng.two_opt_batch(mode, tours[lo:hi], costs[lo:hi])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt, float(st.moves), float(st.evals), float(o.sum())], dtype=torch.float64, device="cuda")
    tmax = t.clone()
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    # parity of the sharded run: rank 0 pushes the WHOLE population through its own GPU and compares every shard
    import hashlib
    mine = hashlib.sha256(np.ascontiguousarray(s).tobytes() + np.ascontiguousarray(o).tobytes()).hexdigest()
    hashes = [None] * world
    if world > 1:
        dist.all_gather_object(hashes, (lo, hi, mine))
    else:
        hashes = [(lo, hi, mine)]
    if rank == 0:
        s1, o1, _ = eng.two_opt_batch(mode, tours, costs)
        ok = all(hashlib.sha256(np.ascontiguousarray(s1[a:b]).tobytes() + np.ascontiguousarray(o1[a:b]).tobytes()).hexdigest() == h
                 for a, b, h in hashes)
        print(json.dumps({"workload": f"GA population: {B} random tours of uni{n}, {'BI' if mode == BI else 'FI'} 2-opt to the local optimum",
                          "n_gpus": world, "seconds": float(tmax[0]), "tours_per_s": B / float(tmax[0]), "moves": int(t[1]),
                          "evals": int(t[2]), "cost_sum": float(t[3]), "shards_equal_single_gpu": bool(ok)}), flush=True)
        if not ok:
            sys.exit(1)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
