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


def run_sharded(sampler, measurement, generators, rank=None, world=None, gather=False, **call_kwargs):
    """Run a `BatchedGuidedSampler` on this rank's clips (clip i -> rank i mod W; the guided path has no collective).

    measurement: (B, ...) one row per clip, or (1, ...) shared; generators: list of B per-clip generators (every rank
    builds the same list and uses its own entries, so a clip's stream does not depend on the world size); the noise
    predictor's `clips` argument carries the GLOBAL clip indices.
    Returns (clip indices of this rank, BatchedSamplerOutput of those clips).  gather=True additionally all-gathers the
    per-clip losses and restart counts (logging only; a few bytes) and returns them as a third item
    {clip index: (loss, restarts)}."""
    import torch.distributed as dist
    if rank is None or world is None:
        on = dist.is_available() and dist.is_initialized()
        rank, world = (dist.get_rank(), dist.get_world_size()) if on else (0, 1)
    B = len(generators)
    ids = shard_indices(B, rank, world)
    out = None
    if ids:
        meas = measurement if measurement.shape[0] == 1 else measurement[ids].contiguous()
        out = sampler(meas, [generators[i] for i in ids], clip_ids=ids, **call_kwargs)
    if not gather:
        return ids, out
    mine = {i: (float(out.loss[k]), int(out.restarts[k])) for k, i in enumerate(ids)} if ids else {}
    if world > 1:
        parts = [None] * world
        dist.all_gather_object(parts, mine)
        mine = {k: v for part in parts for k, v in part.items()}
    return ids, out, mine
