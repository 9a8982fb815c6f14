"""GPU test of the sharded path on >= 2 GPUs (skipped on a 1-GPU box): tools/mgpu_check.py under torchrun."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_sharded_bi_equals_single_gpu():
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "tools", "mgpu_check.py"), "6000", "60"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_sharded_exact_scan_geo_equals_single_gpu():
    """GEO (gr666) shards the exact FP64 scan over the ranks; the argmin key goes through the NCCL min-allreduce."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29519", os.path.join(ROOT, "tools", "mgpu_check.py"), "0", "200", "gr666"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_sharded_tour_batches_equal_single_gpu():
    """independent tour batches (GA population, multi-start): contiguous shards per rank, no collective on the data path;
    every shard's tours and costs must equal the same tours pushed through one GPU."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29518", os.path.join(ROOT, "tools", "batch_mgpu.py"), "FI"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and '"shards_equal_single_gpu": true' in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
