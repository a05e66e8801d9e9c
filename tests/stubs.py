"""Deterministic, differentiable stand-ins for the networks that stay in PyTorch (AutoencoderKL decoder and
SpeechT5HifiGan vocoder) plus the synthetic inputs of SURVEY.md section 8(d).  Shared by the golden generator,
the CPU oracle tests, the GPU parity tests, smoke() and bench.py so every side sees identical inputs.

Shapes follow pipeline_musicldm.py:406-412,602-609: latent (B, 8, H, 16) -> mel (B, 1, 4H, 64) -> wav (B, 4H*160 + 32).
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import torch

SR = 16000
HOP = 160


class StubVAE(torch.nn.Module):
    """`vae.decode(z).sample` and `vae.config.scaling_factor` (scheduling_dps.py:195-197)."""

    def __init__(self, seed=11, scaling_factor=0.18215):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.weight = torch.nn.Parameter(torch.randn(8, 1, 4, 4, generator=g) * 0.35, requires_grad=False)
        self.config = SimpleNamespace(scaling_factor=scaling_factor)

    def decode(self, z):
        mel = torch.nn.functional.conv_transpose2d(z, self.weight.to(z.dtype), stride=4)
        return SimpleNamespace(sample=4.0 * torch.tanh(0.25 * mel) - 5.0)


class StubVocoder(torch.nn.Module):
    """`vocoder(mel)` : (B, T, 64) -> (B, T*160 + 32) (SpeechT5HifiGan emits a little more than L samples)."""

    def __init__(self, seed=12, extra=32):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.proj = torch.nn.Parameter(torch.randn(64, HOP, generator=g) / 8.0, requires_grad=False)
        t = torch.arange(HOP, dtype=torch.float32)
        self.carrier = torch.nn.Parameter(torch.sin(2 * math.pi * 5.0 * t / HOP)[None, None, :] * 0.05,
                                          requires_grad=False)
        self.extra = extra

    def forward(self, mel):
        frames = torch.tanh(mel @ self.proj.to(mel.dtype)) * 0.3 + self.carrier.to(mel.dtype)
        wav = frames.reshape(mel.shape[0], -1)
        return torch.nn.functional.pad(wav, (0, self.extra))


def synth_clip(i, length, kind="music"):
    """clip i of SURVEY.md 8(d): seeded fp32 waveform of `length` samples."""
    g = torch.Generator().manual_seed(1000 + i)
    t = torch.arange(length, dtype=torch.float32) / SR
    noise = torch.randn(length, generator=g)
    if kind == "noise":
        return 0.1 * noise
    return (0.3 * torch.sin(2 * math.pi * 220.0 * t)
            + 0.2 * torch.sin(2 * math.pi * 1760.0 * t) * torch.exp(-t) + 0.02 * noise)


def synth_clips(b, length, first=0, kind="music"):
    return torch.stack([synth_clip(first + i, length, kind) for i in range(b)])


def synth_latents(b, h, first=0):
    """(x_t, eps) per clip, seeded 2000+i; shape (b, 8, h, 16)."""
    xs, es = [], []
    for i in range(b):
        g = torch.Generator().manual_seed(2000 + first + i)
        xs.append(torch.randn(1, 8, h, 16, generator=g))
        es.append(torch.randn(1, 8, h, 16, generator=g))
    return torch.cat(xs), torch.cat(es)


def step_generators(b, first=0, device="cpu"):
    return [torch.Generator(device=device).manual_seed(3000 + first + i) for i in range(b)]


MUSICLDM_SCHED = dict(num_train_timesteps=1000, beta_start=0.0015, beta_end=0.0195, beta_schedule="scaled_linear",
                      trained_betas=None, clip_sample=False, set_alpha_to_one=False, steps_offset=1,
                      prediction_type="epsilon", thresholding=False, dynamic_thresholding_ratio=0.995,
                      clip_sample_range=1.0, sample_max_value=1.0, timestep_spacing="leading",
                      rescale_betas_zero_snr=False)  # configs/model/musicldm.yaml:7-22


# ---- mel_spectrogram_to_waveform_with_phase cases (tests/golden/make_istft_golden.py, istft.npz) ----
#: name -> (B, T, phase shared by the batch, seed, original_waveform_length: 0 = as is / clipped / zero-padded)
ISTFT_CASES = {"b1_t26": (1, 26, True, 11, 0), "b3_t41_own_phase": (3, 41, False, 12, 6000),
               "b2_t9": (2, 9, True, 13, 2000)}


def istft_inputs(name):
    """mel (B, 1, T, 64) in a dB-like range with negative-going frames (so that the relu of InverseMelScale acts) and
    phase (1 or B, 513, T) in [-pi, pi)."""
    B, T, shared, seed, _ = ISTFT_CASES[name]
    g = torch.Generator().manual_seed(seed)
    mel = torch.rand(B, 1, T, 64, generator=g) * 6.0 - 1.0
    phase = (torch.rand(1 if shared else B, 513, T, generator=g) * 2.0 - 1.0) * math.pi
    return mel, phase


#: waveform_to_spectrogram fixtures of istft.npz: name -> clip length (clips synth_clips(2, L, first=70))
SPECTROGRAM_CASES = {"w2s_4000": 4000, "w2s_4133": 4133}


# ---- FAD statistics / Frechet distance cases (tests/golden/make_fad_golden.py, fad.npz) ----
#: name -> (frames of set 1, frames of set 2, embedding width d, files the first set is split into for the online merge)
FAD_CASES = {"d128": (900, 700, 128, 5), "d64_few": (90, 400, 64, 3), "d96": (1500, 1200, 96, 7)}


def fad_embeddings(name):
    """two seeded (n, d) fp16 embedding sets, as fadtk caches them (model_loader.py:46-48): correlated features with
    different means and scales, so that the covariances are neither diagonal nor equal"""
    import numpy as np
    n1, n2, d, _ = FAD_CASES[name]
    rng = np.random.default_rng(100 + d + n1)
    mix1 = rng.standard_normal((d, d)) / np.sqrt(d)
    mix2 = mix1 + 0.3 * rng.standard_normal((d, d)) / np.sqrt(d)
    a = rng.standard_normal((n1, d)) @ mix1 * 1.5 + rng.standard_normal(d) * 0.4
    b = rng.standard_normal((n2, d)) @ mix2 * 1.2 + rng.standard_normal(d) * 0.4 + 0.1
    return a.astype(np.float16), b.astype(np.float16)
