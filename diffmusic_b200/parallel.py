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


class _RawCudaBuffer:
    """`__cuda_array_interface__` view of raw device memory, so torch can alias it (torch.as_tensor)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class PeerGroup:
    """Device buffers of every rank of ONE node mapped into every other rank's address space (CUDA IPC), for kernels
    that read their peers' memory over NVLink directly (the FAD moment exchange, csrc/fad_exchange.cu).

    Each rank owns one zeroed allocation holding `n_doubles` float64 followed by an int32 flag pad (dm_peer_alloc); the
    64-byte IPC handles travel once through `all_gather_object` (host side, start-up only) and every peer's handle is
    opened with THIS rank's device current (dm_ipc_open: lazy peer access), which is what lets a kernel on this device
    dereference the mapping.  `ptrs` / `flag_ptrs` list the device addresses of rank 0..W-1's buffers as seen from this
    process (entry `rank` is the local one); `buf` / `flags` are torch views of the local allocation.  `close()` -- a
    barrier, then unmap and free -- must run before the process group goes away."""

    def __init__(self, n_doubles, flag_words, device=None, group=None):
        import ctypes as C
        import torch.distributed as dist
        from . import _lib
        self.group = group
        on = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if on else 0
        self.world = dist.get_world_size(group) if on else 1
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        dev = self.device.index
        self._flag_off = 8 * int(n_doubles)
        nbytes = self._flag_off + 4 * int(flag_words)
        nbytes = (nbytes + 255) & ~255
        p = C.c_void_p()
        _lib.call("dm_peer_alloc", dev, nbytes, C.byref(p))
        self._base = int(p.value)
        self.buf = torch.as_tensor(_RawCudaBuffer(self._base, (int(n_doubles),), "<f8"), device=self.device)
        self.flags = torch.as_tensor(_RawCudaBuffer(self._base + self._flag_off, (int(flag_words),), "<i4"),
                                     device=self.device)
        bases = [self._base] * self.world
        self._opened = []
        if self.world > 1:
            h = (C.c_ubyte * 64)()
            _lib.call("dm_ipc_export", dev, self._base, h)
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(h), group=group)
            for r, hb in enumerate(handles):
                if r == self.rank:
                    continue
                q = C.c_void_p()
                _lib.call("dm_ipc_open", dev, (C.c_ubyte * 64).from_buffer_copy(hb), C.byref(q))
                bases[r] = int(q.value)
                self._opened.append(bases[r])
            dist.barrier(group=group)
        self.ptrs = list(bases)
        self.flag_ptrs = [b + self._flag_off for b in bases]

    def close(self):
        import torch.distributed as dist
        from . import _lib
        if self._base is None:
            return
        torch.cuda.synchronize(self.device)
        if self.world > 1 and dist.is_initialized():
            dist.barrier(group=self.group)  # nobody is still reading this rank's buffer
        for q in self._opened:
            _lib.call("dm_ipc_close", self.device.index, q)
        self._opened = []
        self.buf = self.flags = None
        _lib.call("dm_peer_free", self.device.index, self._base)
        self._base = None
