"""Multi-GPU plumbing: one process per GPU (torchrun), clips sharded rank::world, no collective on the guided path.

The guided trajectories of different clips are independent, so the only collective on the whole path is
the FAD moment all-reduce (diffmusic_b200/fad.py).
"""
from __future__ import annotations

import os

import torch


def shard_indices(n_items, rank, world):
    """clip i -> rank i mod W."""
    return list(range(rank, n_items, world))


def init_distributed(backend=None):
    """Initialise torch.distributed from the torchrun environment; returns (rank, world, local_rank)."""
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local
