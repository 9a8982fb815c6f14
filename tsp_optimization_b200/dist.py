"""torch.distributed plumbing for the multi-GPU paths (one process per GPU).

PyTorch is used only to bootstrap the ranks: it broadcasts the NCCL unique id that the native engine then
uses for its OWN communicator (csrc/engine.cu dlopens libnccl.so.2 and issues one 8-byte
ncclAllReduce(min) per best-improvement pass on the engine's stream).  Nothing here touches tour data.
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.distributed as dist


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init_process_group_from_env(backend: str | None = None):
    rank, world, local = env_rank_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


def attach_engine_comm(engine, rank: int, world: int):
    """Create the engine's native NCCL communicator: rank 0 makes the id, torch broadcasts it."""
    if world == 1:
        return
    from .engine import Engine
    uid = Engine.comm_unique_id() if rank == 0 else bytes(128)
    t = torch.tensor(list(uid), dtype=torch.uint8)
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.broadcast(t, src=0)
    engine.comm_init(bytes(t.cpu().tolist()), rank, world)


def allreduce_min_key(packed: int) -> int:
    """Min-allreduce of one packed (delta,i,j) key through torch.distributed (CPU/gloo tests and fallback
    bootstrap only; the product path reduces on the device with the engine's own NCCL communicator)."""
    # gloo has no uint64: split into two non-negative int64 halves that keep the ordering
    hi, lo = packed >> 32, packed & 0xFFFFFFFF
    t = torch.tensor([hi], dtype=torch.int64)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    best_hi = int(t.item())
    t2 = torch.tensor([lo if hi == best_hi else (1 << 40)], dtype=torch.int64)
    dist.all_reduce(t2, op=dist.ReduceOp.MIN)
    return (best_hi << 32) | int(t2.item())


def shard_batch(batch: int, rank: int, world: int):
    """Contiguous block of independent tours for this rank (multi-start / GA population sharding)."""
    per = (batch + world - 1) // world
    lo = min(batch, rank * per)
    return lo, min(batch, lo + per)
