"""FAD embedding statistics on the GPU: mean / covariance of (N, d) fp16 embeddings (fadtk/fad.py:41-47) and the
multi-file / multi-rank merge (fadtk/utils.py:13-46).

Every shard (file, clip batch or rank) contributes raw float64 moments  acc = [n | sum x | sum x x^T]  through
dm_fad_moments; shards add, ranks add with ONE all-reduce (NCCL over NVLink when torch.distributed is initialised with
the nccl backend), and  mu = sx / n,  cov = (sxx - n mu mu^T) / (n - 1)  is the Chan merge the reference computes
pairwise on the host.
"""
from __future__ import annotations

import torch

from . import _lib


class EmbeddingMoments:
    """Accumulator of raw moments for one embedding model of width d."""

    ENGINES = {"auto": 0, "simt": 1, "tcgen05": 2}

    def __init__(self, d, device=None, engine="auto"):
        self.d = int(d)
        self.engine = self.ENGINES[engine]
        # the accumulator may live on the CPU (gloo tests of the exchange step); update()/finalize() need CUDA
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.acc = torch.zeros(1 + self.d + self.d * self.d, device=self.device, dtype=torch.float64)

    def update(self, embd):
        """embd: (n_frames, d) fp16 (what fadtk caches, model_loader.py:46-48) or any float dtype (cast to fp16 only
        if it already is fp16-representable is the caller's business: fp32 input is rounded to fp16 like the cache)."""
        if embd.dim() != 2 or embd.shape[1] != self.d:
            raise ValueError(f"expected (n, {self.d}) embeddings, got {tuple(embd.shape)}")
        if self.device.type != "cuda":
            raise _lib.DiffMusicB200Error("dm_fad_moments needs a CUDA accumulator (no CPU fallback)")
        x = embd.to(device=self.device, dtype=torch.float16).contiguous()
        if x.shape[0] == 0:
            return self
        _lib.call("dm_fad_moments_ex", x.data_ptr(), x.shape[0], self.d, self.acc.data_ptr(), self.engine,
                  _lib.stream())
        return self

    def all_reduce(self, group=None):
        """Sum the moments over ranks: one collective of 1 + d + d^2 float64 (the only exchange on this path)."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.acc, op=dist.ReduceOp.SUM, group=group)
        return self

    def count(self):
        return int(round(float(self.acc[0].item())))

    def finalize(self):
        """(mu (d,), cov (d, d)) float64 on the device; cov is zeros when fewer than 2 frames (fadtk/utils.py:42-46)."""
        if self.device.type != "cuda":
            raise _lib.DiffMusicB200Error("dm_fad_finalize needs a CUDA accumulator (no CPU fallback)")
        mu = torch.empty(self.d, device=self.device, dtype=torch.float64)
        cov = torch.empty((self.d, self.d), device=self.device, dtype=torch.float64)
        _lib.call("dm_fad_finalize", self.acc.data_ptr(), self.d, mu.data_ptr(), cov.data_ptr(), _lib.stream())
        return mu, cov


def calc_embd_statistics(embd_lst):
    """fadtk/fad.py:41-47 on the GPU: (mean, cov) of one (n, d) block, returned as float64 NumPy arrays."""
    if embd_lst.shape[0] < 2:
        raise AssertionError(f"FAD requires at least two embedding window frames, you have {tuple(embd_lst.shape)}.")
    t = torch.as_tensor(embd_lst)
    mu, cov = EmbeddingMoments(t.shape[1]).update(t).finalize()
    return mu.cpu().numpy(), cov.cpu().numpy()


def calculate_embd_statistics_online(arrays, group=None):
    """fadtk/utils.py:19-46 with in-memory blocks (or .npy paths) instead of a file list; all-reduced over `group`
    when torch.distributed is initialised, so each rank may pass only its own shard."""
    import numpy as np
    if len(arrays) == 0:
        raise AssertionError("No files provided")
    first = np.load(arrays[0]) if isinstance(arrays[0], (str, bytes)) or hasattr(arrays[0], "__fspath__") else arrays[0]
    mom = EmbeddingMoments(first.shape[-1])
    for a in arrays:
        if isinstance(a, (str, bytes)) or hasattr(a, "__fspath__"):
            a = np.load(a)
        mom.update(torch.as_tensor(a))
    mom.all_reduce(group)
    mu, cov = mom.finalize()
    return mu.cpu().numpy(), cov.cpu().numpy()
