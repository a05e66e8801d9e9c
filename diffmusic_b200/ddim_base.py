"""Base scheduler plumbing: diffusers' DDIMScheduler when it is installed, else a dependency-free restatement.

The guided schedulers only need from the base: the beta / alphas_cumprod tables, final_alpha_cumprod,
set_timesteps, timesteps, _get_variance, init_noise_sigma, scale_model_input, order and `.config`.  The arithmetic of
`step` itself lives in csrc/sched_update.cu.  The restatement follows diffusers==0.31.0 (requirements.txt:6 of the
reference); diffusers is not vendored by the reference and no reference test pins this boundary (parity unpinned).
"""
from __future__ import annotations

import functools
import inspect
import math
from dataclasses import fields
from types import SimpleNamespace

import numpy as np
import torch

try:  # the deployment the reference pipelines run in
    from diffusers.configuration_utils import register_to_config  # type: ignore
    from diffusers.schedulers import DDIMScheduler as DDIMBase  # type: ignore
    from diffusers.utils import BaseOutput  # type: ignore
    HAVE_DIFFUSERS = True
except Exception:  # diffusers absent (this build container): stand-alone base
    HAVE_DIFFUSERS = False

    def register_to_config(init):
        """Record the constructor arguments on `self.config` (what diffusers' decorator of the same name does)."""
        sig = inspect.signature(init)

        @functools.wraps(init)
        def wrapper(self, *args, **kwargs):
            init(self, *args, **kwargs)
            bound = sig.bind_partial(self, *args, **kwargs)
            cfg = getattr(self, "config", None) or SimpleNamespace()
            for name, p in sig.parameters.items():
                if name == "self" or p.kind in (p.VAR_POSITIONAL, p.VAR_KEYWORD):
                    continue
                setattr(cfg, name, bound.arguments.get(name, p.default))
            self.config = cfg

        return wrapper

    class BaseOutput:
        """dataclass with dict-style access, like diffusers.utils.BaseOutput."""

        def __getitem__(self, key):
            if isinstance(key, str):
                return getattr(self, key)
            return self.to_tuple()[key]

        def keys(self):
            return [f.name for f in fields(self) if getattr(self, f.name) is not None]

        def to_tuple(self):
            return tuple(getattr(self, k) for k in self.keys())

    def _cosine_betas(n, max_beta=0.999):
        bar = lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2  # noqa: E731
        return torch.tensor([min(1 - bar((i + 1) / n) / bar(i / n), max_beta) for i in range(n)],
                            dtype=torch.float32)

    def _zero_terminal_snr(betas):
        s = torch.cumprod(1.0 - betas, dim=0).sqrt()
        first, last = s[0].clone(), s[-1].clone()
        s = (s - last) * (first / (first - last))
        bar = s ** 2
        return 1 - torch.cat([bar[0:1], bar[1:] / bar[:-1]])

    class DDIMBase:
        """Tables and timestep bookkeeping of diffusers 0.31.0 DDIMScheduler (no `step`)."""

        order = 1

        def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear",
                     trained_betas=None, clip_sample=True, set_alpha_to_one=True, steps_offset=0,
                     prediction_type="epsilon", thresholding=False, dynamic_thresholding_ratio=0.995,
                     clip_sample_range=1.0, sample_max_value=1.0, timestep_spacing="leading",
                     rescale_betas_zero_snr=False):
            self.config = SimpleNamespace(
                num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
                beta_schedule=beta_schedule, trained_betas=trained_betas, clip_sample=clip_sample,
                set_alpha_to_one=set_alpha_to_one, steps_offset=steps_offset, prediction_type=prediction_type,
                thresholding=thresholding, dynamic_thresholding_ratio=dynamic_thresholding_ratio,
                clip_sample_range=clip_sample_range, sample_max_value=sample_max_value,
                timestep_spacing=timestep_spacing, rescale_betas_zero_snr=rescale_betas_zero_snr)
            f32 = torch.float32
            if trained_betas is not None:
                betas = torch.tensor(trained_betas, dtype=f32)
            elif beta_schedule == "linear":
                betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=f32)
            elif beta_schedule == "scaled_linear":
                betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=f32) ** 2
            elif beta_schedule == "squaredcos_cap_v2":
                betas = _cosine_betas(num_train_timesteps)
            else:
                raise NotImplementedError(f"{beta_schedule} is not implemented for {self.__class__}")
            if rescale_betas_zero_snr:
                betas = _zero_terminal_snr(betas)
            self.betas = betas
            self.alphas = 1.0 - betas
            self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
            self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
            self.init_noise_sigma = 1.0
            self.num_inference_steps = None
            self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))

        def scale_model_input(self, sample, timestep=None):
            return sample

        def _get_variance(self, timestep, prev_timestep):
            a_t = self.alphas_cumprod[timestep]
            a_p = self.alphas_cumprod[prev_timestep] if prev_timestep >= 0 else self.final_alpha_cumprod
            return ((1 - a_p) / (1 - a_t)) * (1 - a_t / a_p)

        def set_timesteps(self, num_inference_steps, device=None):
            n_train = self.config.num_train_timesteps
            if num_inference_steps > n_train:
                raise ValueError(f"`num_inference_steps`: {num_inference_steps} cannot be larger than "
                                 f"`self.config.train_timesteps`: {n_train}")
            self.num_inference_steps = num_inference_steps
            mode = self.config.timestep_spacing
            if mode == "linspace":
                ts = np.linspace(0, n_train - 1, num_inference_steps).round()[::-1].copy().astype(np.int64)
            elif mode == "leading":
                ts = (np.arange(0, num_inference_steps) * (n_train // num_inference_steps)).round()[::-1]
                ts = ts.copy().astype(np.int64) + self.config.steps_offset
            elif mode == "trailing":
                ts = np.round(np.arange(n_train, 0, -n_train / num_inference_steps)).astype(np.int64) - 1
            else:
                raise ValueError(f"{mode} is not supported. Choose one of 'leading', 'trailing' or 'linspace'.")
            self.timesteps = torch.from_numpy(ts).to(device)


def _clip_generators(shape, generator, device, dtype, layout):
    """The list of per-clip CUDA generators when the batched Philox kernel (csrc/rng_clips.cu) can stand in for the
    reference's per-clip torch.randn loop, else None."""
    from . import _lib
    if not isinstance(generator, (list, tuple)) or not 2 <= len(generator) <= 128 or len(generator) != shape[0]:
        return None
    if device.type != "cuda" or (layout not in (None, torch.strided)):
        return None
    if device.index is not None and device.index != torch.cuda.current_device():
        return None  # the kernel launches on the current device (one process per GPU is the deployment model)
    if (dtype or torch.get_default_dtype()) not in _lib.IO_DTYPES:
        return None
    for g in generator:
        gd = g.device
        if gd.type != "cuda" or (gd.index is not None and device.index is not None and gd.index != device.index):
            return None
    return generator


_INC_CACHE = {}


def _offset_increment(n, device):
    """how far one draw of n values advances a generator's Philox offset (torch's launch geometry, per device model)"""
    from . import _lib
    inc = _INC_CACHE.get(n)
    if inc is None:
        with torch.cuda.device(device):
            inc = _INC_CACHE[n] = int(_lib.load().dm_randn_offset_increment(n))
    return inc


def _clip_numel(shape):
    n = 1
    for d in shape[1:]:
        n *= int(d)
    return n


def skip_randn(shape, generator, device=None, dtype=None):
    """Consume exactly what `randn_tensor(shape, generator, ...)` would consume without producing the values (the
    reference's discarded base-step draw, scheduling_dps.py:166-175): for per-clip CUDA generators that is an offset
    bump, otherwise the draw itself."""
    device = torch.device(device) if device is not None else torch.device("cpu")
    gens = _clip_generators(shape, generator, device, dtype, None)
    if gens is None:
        randn_tensor(shape, generator=generator, device=device, dtype=dtype)
        return
    inc = _offset_increment(_clip_numel(shape), device)
    for g in gens:
        g.set_offset(g.get_offset() + inc)


def randn_clips_f32(shape, generator, device=None, dtype=None, out=None):
    """fp32 tensor holding the values of `randn_tensor(shape, generator=[g_0..g_{B-1}], dtype=dtype)` (rounded through
    `dtype` when it is 16-bit), drawn for all clips by ONE kernel with torch's own Philox / curand_normal4 mapping and
    each generator's (seed, offset); the generators' offsets advance as if torch had drawn.  `out`: optional contiguous
    fp32 destination of that shape.  None if the fused path does not apply (single generator, CPU generators, other
    dtypes)."""
    import ctypes as C
    from . import _lib
    device = torch.device(device) if device is not None else torch.device("cpu")
    gens = _clip_generators(shape, generator, device, dtype, None)
    if gens is None:
        return None
    B, n = len(gens), _clip_numel(shape)
    offsets = [g.get_offset() for g in gens]
    seeds = (C.c_ulonglong * B)(*[g.initial_seed() for g in gens])
    offs = (C.c_ulonglong * B)(*offsets)
    if out is None:
        out = torch.empty(tuple(shape), device=device, dtype=torch.float32)
    elif out.dtype != torch.float32 or not out.is_contiguous() or out.numel() != B * n or out.device != device:
        raise ValueError("randn_clips_f32: `out` must be a contiguous fp32 tensor of the requested shape")
    _lib.call("dm_randn_clips", seeds, offs, B, n, _lib.IO_DTYPES[dtype or torch.get_default_dtype()], out.data_ptr(),
              _lib.stream(device))
    inc = _offset_increment(n, device)
    for g, o in zip(gens, offsets):
        g.set_offset(o + inc)
    return out.view(tuple(shape))


def randn_tensor(shape, generator=None, device=None, dtype=None, layout=None):
    """Noise draw with the reference's generator semantics (diffmusic/torch_utils.py:31-76): a single generator draws
    the whole batch, a list draws one (1, ...) tensor per clip; CPU generators draw on the CPU and the result is moved.
    Single generators draw through torch; a list of CUDA generators goes through one batched kernel that reproduces
    torch's Philox stream for every clip bit for bit (csrc/rng_clips.cu, tests/test_gpu_parity.py)."""
    device = torch.device(device) if device is not None else torch.device("cpu")
    layout = layout or torch.strided
    where = device
    if generator is not None:
        first = generator[0] if isinstance(generator, (list, tuple)) else generator
        kind = first.device.type
        if kind != device.type and kind == "cpu":
            where = torch.device("cpu")
        elif kind != device.type and kind == "cuda":
            raise ValueError(f"Cannot generate a {device} tensor from a generator of type {kind}.")
    if isinstance(generator, (list, tuple)) and len(generator) == 1:
        generator = generator[0]
    if isinstance(generator, (list, tuple)):
        fused = randn_clips_f32(shape, generator, where, dtype) if where.type == "cuda" and layout == torch.strided \
            else None
        if fused is not None:  # one launch for all clips, same values and generator states as the loop below
            return fused.to(dtype or torch.get_default_dtype()).to(device)
        per = (1,) + tuple(shape[1:])
        out = torch.cat([torch.randn(per, generator=generator[i], device=where, dtype=dtype, layout=layout)
                         for i in range(shape[0])], dim=0)
        return out.to(device)
    return torch.randn(tuple(shape), generator=generator, device=where, dtype=dtype, layout=layout).to(device)
