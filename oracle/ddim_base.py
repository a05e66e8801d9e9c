"""Restatement of diffusers==0.31.0 ``DDIMScheduler`` (the base class of every reference scheduler,
diffmusic/schedulers/scheduling_dps.py:5-7,44-60).  TEST INFRASTRUCTURE -- see oracle/__init__.py.

diffusers is absent from this image and from /root/reference, so this follows the published 0.31.0 source
(`src/diffusers/schedulers/scheduling_ddim.py`).  PARITY UNPINNED: no reference test pins this boundary.

Reference call sites this must serve:
  super().__init__(...)                     scheduling_dps.py:44-60
  super().step(...).pred_original_sample    scheduling_ddim.py:84-93, scheduling_dps.py:166-175,
                                            scheduling_mpgd.py:164-173, scheduling_dsg.py:178-186,
                                            scheduling_diffmusic.py:180-188
  self._get_variance(t, t_prev)             scheduling_dps.py:161
  self.alphas_cumprod / final_alpha_cumprod scheduling_dps.py:158-160
  set_timesteps / timesteps / init_noise_sigma / scale_model_input / order   pipeline_musicldm.py:655-693
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np
import torch


class DDIMOutput(SimpleNamespace):
    """prev_sample / pred_original_sample holder (diffusers DDIMSchedulerOutput)."""


def _betas_for_alpha_bar(n, max_beta=0.999):
    def bar(t):
        return math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2

    betas = [min(1 - bar((i + 1) / n) / bar(i / n), max_beta) for i in range(n)]
    return torch.tensor(betas, dtype=torch.float32)


def _rescale_zero_terminal_snr(betas):
    alphas = 1.0 - betas
    abar_sqrt = torch.cumprod(alphas, dim=0).sqrt()
    a0 = abar_sqrt[0].clone()
    aT = abar_sqrt[-1].clone()
    abar_sqrt = (abar_sqrt - aT) * (a0 / (a0 - aT))
    abar = abar_sqrt ** 2
    alphas = torch.cat([abar[0:1], abar[1:] / abar[:-1]])
    return 1 - alphas


class DDIMSchedulerBase:
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
        if trained_betas is not None:
            self.betas = torch.tensor(trained_betas, dtype=torch.float32)
        elif beta_schedule == "linear":
            self.betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        elif beta_schedule == "scaled_linear":
            self.betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps,
                                        dtype=torch.float32) ** 2
        elif beta_schedule == "squaredcos_cap_v2":
            self.betas = _betas_for_alpha_bar(num_train_timesteps)
        else:
            raise NotImplementedError(f"{beta_schedule} is not implemented for {self.__class__}")
        if rescale_betas_zero_snr:
            self.betas = _rescale_zero_terminal_snr(self.betas)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))

    def scale_model_input(self, sample, timestep=None):
        return sample

    def _get_variance(self, timestep, prev_timestep):
        a_t = self.alphas_cumprod[timestep]
        a_prev = self.alphas_cumprod[prev_timestep] if prev_timestep >= 0 else self.final_alpha_cumprod
        return ((1 - a_prev) / (1 - a_t)) * (1 - a_t / a_prev)

    def set_timesteps(self, num_inference_steps, device=None):
        T = self.config.num_train_timesteps
        if num_inference_steps > T:
            raise ValueError("num_inference_steps cannot exceed num_train_timesteps")
        self.num_inference_steps = num_inference_steps
        spacing = self.config.timestep_spacing
        if spacing == "linspace":
            ts = np.linspace(0, T - 1, num_inference_steps).round()[::-1].copy().astype(np.int64)
        elif spacing == "leading":
            ratio = T // num_inference_steps
            ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64)
            ts += self.config.steps_offset
        elif spacing == "trailing":
            ratio = T / num_inference_steps
            ts = np.round(np.arange(T, 0, -ratio)).astype(np.int64)
            ts -= 1
        else:
            raise ValueError(f"{spacing} is not supported")
        self.timesteps = torch.from_numpy(ts).to(device)

    def _threshold_sample(self, sample):
        dtype = sample.dtype
        b = sample.shape[0]
        flat = sample.float().reshape(b, -1)
        s = torch.quantile(flat.abs(), self.config.dynamic_thresholding_ratio, dim=1)
        s = torch.clamp(s, min=1, max=self.config.sample_max_value).unsqueeze(1)
        flat = torch.clamp(flat, -s, s) / s
        return flat.reshape(sample.shape).to(dtype)

    def step(self, model_output, timestep, sample, eta=0.0, use_clipped_model_output=False, generator=None,
             variance_noise=None, return_dict=True):
        if self.num_inference_steps is None:
            raise ValueError("run set_timesteps first")
        prev_timestep = timestep - self.config.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[timestep]
        a_prev = self.alphas_cumprod[prev_timestep] if prev_timestep >= 0 else self.final_alpha_cumprod
        b_t = 1 - a_t
        ptype = self.config.prediction_type
        if ptype == "epsilon":
            x0 = (sample - b_t ** 0.5 * model_output) / a_t ** 0.5
            eps = model_output
        elif ptype == "sample":
            x0 = model_output
            eps = (sample - a_t ** 0.5 * x0) / b_t ** 0.5
        elif ptype == "v_prediction":
            x0 = (a_t ** 0.5) * sample - (b_t ** 0.5) * model_output
            eps = (a_t ** 0.5) * model_output + (b_t ** 0.5) * sample
        else:
            raise ValueError(f"unknown prediction_type {ptype}")
        if self.config.thresholding:
            x0 = self._threshold_sample(x0)
        elif self.config.clip_sample:
            x0 = x0.clamp(-self.config.clip_sample_range, self.config.clip_sample_range)
        variance = self._get_variance(timestep, prev_timestep)
        std = eta * variance ** 0.5
        if use_clipped_model_output:
            eps = (sample - a_t ** 0.5 * x0) / b_t ** 0.5
        prev = a_prev ** 0.5 * x0 + (1 - a_prev - std ** 2) ** 0.5 * eps
        if eta > 0:
            if variance_noise is not None and generator is not None:
                raise ValueError("Cannot pass both generator and variance_noise.")
            if variance_noise is None:
                variance_noise = randn_like_reference(model_output.shape, generator, model_output.device,
                                                      model_output.dtype)
            prev = prev + std * variance_noise
        if not return_dict:
            return (prev, x0)
        return DDIMOutput(prev_sample=prev, pred_original_sample=x0)


def randn_like_reference(shape, generator, device, dtype):
    """diffmusic/torch_utils.py:31-76 (== diffusers.utils.torch_utils.randn_tensor): one draw of `shape`, or per-sample
    draws of (1, ...) concatenated when a list of generators is given; CPU generators draw on CPU then move."""
    device = torch.device(device) if device is not None else torch.device("cpu")
    rand_device = device
    if generator is not None:
        g0 = generator[0] if isinstance(generator, (list, tuple)) else generator
        if g0.device.type != device.type and g0.device.type == "cpu":
            rand_device = torch.device("cpu")
        elif g0.device.type != device.type and g0.device.type == "cuda":
            raise ValueError(f"Cannot generate a {device} tensor from a generator of type cuda.")
    if isinstance(generator, (list, tuple)) and len(generator) == 1:
        generator = generator[0]
    if isinstance(generator, (list, tuple)):
        one = (1,) + tuple(shape[1:])
        parts = [torch.randn(one, generator=generator[i], device=rand_device, dtype=dtype)
                 for i in range(shape[0])]
        return torch.cat(parts, dim=0).to(device)
    return torch.randn(tuple(shape), generator=generator, device=rand_device, dtype=dtype).to(device)
