"""CPU test of the multi-GPU neighbourhood sharding (world_size 2, gloo): each rank evaluates ONLY the pairs
of its round-robin share of the tile plan (here with the oracle's distances, in position space like the
kernel), packs its best key exactly like the device code, and one 8-byte min-allreduce selects the move.
The result must be the reference's best-improvement move, on every rank."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    from oracle.oracle import Oracle
    from tsp_optimization_b200.dist import allreduce_min_key
    from tsp_optimization_b200.engine import key_pack, key_unpack, tile_plan
    from tsp_optimization_b200.instances import succ_to_order, uniform_instance
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc = Oracle()
    n = 600
    xy = uniform_instance(n)
    succ, _ = orc.nn_tour(xy, 0, 0)
    moves = []
    for step in range(3):
        order = succ_to_order(succ)
        D = orc.dist_matrix(xy, 0).astype(np.int64)
        T, R, TJ, rs, rj = tile_plan(n, 2, 64, threads=64)
        TI = T * R
        best = key_pack(0, 0x1FFFF, 0x1FFFF)
        for t in range(rank, int(rs[-1]), world):
            I = int(np.searchsorted(rs, t, side="right") - 1)
            J = int(rj[I] + (t - rs[I]))
            for p in range(I * TI, min((I + 1) * TI, n)):
                for q in range(max(J * TJ, p + 2), min((J + 1) * TJ, n)):
                    if p == 0 and q == n - 1:
                        continue
                    u, v, u1, v1 = order[p], order[q], order[(p + 1) % n], order[(q + 1) % n]
                    delta = D[u, v] + D[u1, v1] - D[u, u1] - D[v, v1]
                    if delta < 0:
                        best = min(best, key_pack(int(delta), int(min(u, v)), int(max(u, v))))
        win = allreduce_min_key(best)
        delta, i, j = key_unpack(win)
        moves.append((i, j, delta))
        # every rank applies the same move to its replica (reference heuristics.c:479-483)
        prev = np.empty(n, dtype=np.int32)
        prev[succ] = np.arange(n, dtype=np.int32)
        a1, b1 = int(succ[i]), int(succ[j])
        succ[i] = j
        succ[a1] = b1
        orc.L.orc_reverse_path(n, succ, j, a1, prev)
    np.save(os.path.join(out_dir, f"moves_{rank}.npy"), np.array(moves, dtype=np.int64))
    dist.destroy_process_group()


def test_sharded_argmin_equals_reference_moves(tmp_path, oracle):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    from tsp_optimization_b200.instances import uniform_instance
    xy = uniform_instance(600)
    succ, _ = oracle.nn_tour(xy, 0, 0)
    _, _, _, log = oracle.two_opt_bi(xy, 0, succ, max_passes=3, log_cap=3)
    m0 = np.load(tmp_path / "moves_0.npy")
    m1 = np.load(tmp_path / "moves_1.npy")
    assert (m0 == m1).all()
    assert m0.tolist() == log.tolist()


def _fi_worker(rank, world, port, out_dir, min_gap=0):
    """first improvement, sharded like csrc/kernels_fi.cu: the row-major pair order from the cursor is cut into segments of SEG
    pairs, rank r scans the segments r, r + world, ... and stops at ITS first improving pair; the min over the ranks of the
    linear index i * n + j is the reference's next move.  min_gap > 0 is the adaptive rule (Ctl::fi_shard): a search is
    sharded (and exchanged) only if the previous one had to sweep more than min_gap pairs — a number every rank has —,
    otherwise every rank searches the whole range alone and NO collective is entered (all ranks must agree on that)."""
    import sys
    sys.path.insert(0, ROOT)
    from oracle.oracle import Oracle
    from tsp_optimization_b200.instances import uniform_instance
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc = Oracle()
    n, SEG = 220, 512
    xy = uniform_instance(n)
    D = orc.dist_matrix(xy, 0).astype(np.int64)
    succ, _ = orc.nn_tour(xy, 0, 0)
    succ = succ.copy()
    pairs = [(i, j) for i in range(n - 1) for j in range(i + 1, n)]  # the reference's enumeration (heuristics.c:466-467)
    moves = []
    cursor, sweep_moves = 0, 0
    NONE = 1 << 40
    shard = min_gap == 0
    exchanges = 0
    while len(moves) < 60:
        found = NONE
        seg = rank if shard else 0
        step = world if shard else 1
        while cursor + seg * SEG < len(pairs):
            lo = cursor + seg * SEG
            for k in range(lo, min(lo + SEG, len(pairs))):
                a, b = pairs[k]
                a1, b1 = int(succ[a]), int(succ[b])
                if b == a1 or b1 == a or a1 == b1:
                    continue
                if D[a, b] + D[a1, b1] - D[a, a1] - D[b, b1] < 0:
                    found = a * n + b
                    break
            if found != NONE:
                break
            seg += step
        f = found
        if shard:
            t = torch.tensor([found], dtype=torch.int64)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            f = int(t.item())
            exchanges += 1
        if f == NONE:  # end of the sweep (heuristics.c:492-496)
            if min_gap:
                shard = len(pairs) - cursor > min_gap
            if sweep_moves == 0:
                break
            cursor, sweep_moves = 0, 0
            continue
        if min_gap:  # the next search's mode, from the length of this one (csrc/tsp_state.cuh fi_advance)
            shard = pairs.index(divmod(f, n)) - cursor + 1 > min_gap
        i, j = divmod(f, n)
        a1, b1 = int(succ[i]), int(succ[j])
        moves.append((i, j, int(D[i, j] + D[a1, b1] - D[i, a1] - D[j, b1])))
        prev = np.empty(n, dtype=np.int32)
        prev[succ] = np.arange(n, dtype=np.int32)
        succ[i] = j
        succ[a1] = b1
        orc.L.orc_reverse_path(n, succ, j, a1, prev)
        sweep_moves += 1
        cursor = pairs.index((i, j)) + 1
    np.save(os.path.join(out_dir, f"fi_moves_{rank}_{min_gap}.npy"), np.array(moves + [(exchanges, 0, 0)], dtype=np.int64))
    dist.destroy_process_group()


def test_sharded_first_improvement_equals_reference_moves(tmp_path, oracle):
    world = 2
    port = _free_port()
    mp.spawn(_fi_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    from tsp_optimization_b200.instances import uniform_instance
    xy = uniform_instance(220)
    succ, cost = oracle.nn_tour(xy, 0, 0)
    _, _, _, log = oracle.two_opt_fi(xy, 0, succ, cost, max_moves=60, log_cap=60)
    m0 = np.load(tmp_path / "fi_moves_0_0.npy")
    m1 = np.load(tmp_path / "fi_moves_1_0.npy")
    assert (m0 == m1).all()
    assert m0[:-1].tolist() == log.tolist()
    # adaptive sharding: same moves, the ranks agree on when to exchange, and they exchange less often than every search
    port = _free_port()
    mp.spawn(_fi_worker, args=(world, port, str(tmp_path), 300), nprocs=world, join=True)
    a0 = np.load(tmp_path / "fi_moves_0_300.npy")
    a1 = np.load(tmp_path / "fi_moves_1_300.npy")
    assert (a0 == a1).all() and a0[:-1].tolist() == log.tolist()
    assert 0 < a0[-1][0] < m0[-1][0]
