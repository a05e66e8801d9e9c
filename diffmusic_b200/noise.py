"""Noise models -- mirror of diffmusic/inverse_problem/noise.py.

GaussianNoise keeps the reference's RNG behaviour (noise.py:13-18: one torch.randn_like draw from the global generator
of the data's device); when it is attached to an operator, the operator adds the drawn noise with a CUDA kernel
(dm_add_scaled).  PoissonNoise is the reference's host-side NumPy round trip (noise.py:21-39): non-differentiable, not in
any shipped config, out of scope for the GPU path and kept only so get_noiser("poisson") keeps working.
"""
from __future__ import annotations

import numpy as np
import torch


class BaseNoise:
    def __call__(self, data):
        return self.forward(data)

    def forward(self, data):
        pass


class GaussianNoise(BaseNoise):
    def __init__(self, sigma):
        self.sigma = sigma

    def forward(self, data):
        noise = torch.randn_like(data, device=data.device)
        if self.sigma == 0 or not data.is_cuda:
            return data + noise * self.sigma
        from . import _lib
        out = data.float().contiguous().clone()
        noise = noise.float().contiguous()
        _lib.call("dm_add_scaled", out.data_ptr(), noise.data_ptr(), float(self.sigma), out.numel(), _lib.stream())
        return out.to(data.dtype)


class PoissonNoise(BaseNoise):
    def __init__(self, rate):
        self.rate = rate

    def forward(self, data):
        x = ((data + 1.0) / 2.0).clamp(0, 1)
        device = x.device
        x = x.detach().cpu()
        x = torch.from_numpy(np.random.poisson(x * 255.0 * self.rate) / 255.0 / self.rate)
        return (x * 2.0 - 1.0).clamp(-1, 1).to(device)


def get_noiser(name, sigma):
    """diffmusic/inverse_problem/__init__.py:4-11."""
    if name == "gaussian":
        return GaussianNoise(sigma)
    if name == "poisson":
        return PoissonNoise(sigma)
    raise ValueError(f"Unknown noise: {name}")
