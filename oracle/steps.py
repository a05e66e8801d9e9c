"""CPU restatement of the five guided `.step` rules -- TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows diffmusic/schedulers/scheduling_ddim.py:78-104, scheduling_dps.py:157-219, scheduling_mpgd.py:157-224,
scheduling_dsg.py:169-230 and scheduling_diffmusic.py:59-68,171-229 with torch autograd on CPU, the restated
diffusers base (oracle/ddim_base.py) underneath.  Pinned against tests/golden/steps.npz (outputs of the reference's
unmodified scheduler files, see tests/golden/make_golden.py).

`reference_step` is the reference verbatim in meaning: every norm spans the whole batch (scheduling_dps.py:211).
`per_clip_step` is the batched semantics the product implements: B independent batch-1
reference trajectories, each with its own generator.
"""
from __future__ import annotations

from types import SimpleNamespace

import torch

from .ddim_base import DDIMSchedulerBase, randn_like_reference

DEFAULTS = {  # (eta, ip_guidance_rate) defaults of each class's .step signature
    "ddim": (0.0, None), "dps": (0.0, 5e-4), "mpgd": (0.0, 1.0), "dsg": (1.0, 0.08), "diffmusic": (0.0, 0.08)}


def make_base(**sched_kwargs):
    return DDIMSchedulerBase(**sched_kwargs)


def _coeffs(base, timestep, eta):
    """scheduling_dps.py:157-162 (identical prologue in all five files); 0-d fp32 CPU tensors."""
    t_prev = timestep - base.config.num_train_timesteps // base.num_inference_steps
    a_t = base.alphas_cumprod[timestep]
    a_prev = base.alphas_cumprod[t_prev] if t_prev >= 0 else base.final_alpha_cumprod
    var = base._get_variance(timestep, t_prev)
    return a_t, 1 - a_t, a_prev, eta * var ** 0.5


def _residual_norm(operator, x0, vae, vocoder, L, measurement, supervised_space):
    """scheduling_dps.py:195-211."""
    mel = vae.decode(1 / vae.config.scaling_factor * x0).sample
    wav = operator.inverse_transform(mel, vocoder)[:, :L]
    pred = operator.forward(wav)
    if supervised_space == "wav_form":
        diff = measurement - pred
    elif supervised_space == "mel_spectrogram":
        diff = operator.transform(measurement) - operator.transform(pred)
    else:
        raise ValueError("supervised_space should be either 'wav_form' or 'mel_spectrogram")
    return torch.linalg.norm(diff)


def slerp(x0, x1, gamma, threshold=0.9995):
    """scheduling_diffmusic.py:59-68."""
    c = ((x0 / torch.norm(x0)) * (x1 / torch.norm(x1))).sum()
    if c.abs() > threshold:
        return x0 + gamma * (x1 - x0)
    th = torch.acos(c)
    s = torch.sin(th)
    return torch.sin((1 - gamma) * th) / s * x0 + torch.sin(gamma * th) / s * x1


def reference_step(kind, base, operator, model_output, timestep, sample, eta=None, ip_guidance_rate=None,
                   generator=None, variance_noise=None, measurement=None, vae=None, vocoder=None,
                   original_waveform_length=0, supervised_space="mel_spectrogram", eps=1e-8):
    d_eta, d_rate = DEFAULTS[kind]
    eta = d_eta if eta is None else eta
    rate = d_rate if ip_guidance_rate is None else ip_guidance_rate
    a_t, b_t, a_prev, std = _coeffs(base, timestep, eta)
    L = original_waveform_length

    def base_x0(x, with_eta=True):
        kw = dict(eta=eta) if with_eta else {}
        return base.step(model_output, timestep, x, generator=generator, variance_noise=variance_noise,
                         **kw).pred_original_sample

    def add_noise(prev):
        if eta > 0:
            if variance_noise is not None and generator is not None:
                raise ValueError("Cannot pass both generator and variance_noise.")
            z = variance_noise if variance_noise is not None else randn_like_reference(
                model_output.shape, generator, model_output.device, model_output.dtype)
            prev = prev + std * z
        return prev

    if kind == "ddim":  # scheduling_ddim.py:78-104
        x0 = base_x0(sample)
        e = (sample - a_t ** 0.5 * x0) / b_t ** 0.5
        prev = a_prev ** 0.5 * x0 + (1 - a_prev) ** 0.5 * e
        return SimpleNamespace(prev_sample=prev.detach(), pred_original_sample=x0, loss=torch.tensor([timestep]))

    if kind == "dps":  # scheduling_dps.py:164-213
        with torch.enable_grad():
            x = sample.clone().detach().requires_grad_(True)
            x0 = base_x0(x)
            e = (x - a_t ** 0.5 * x0) / b_t ** 0.5
            prev = add_noise(a_prev ** 0.5 * x0 + (1 - a_prev - std ** 2) ** 0.5 * e)
            loss = _residual_norm(operator, x0, vae, vocoder, L, measurement, supervised_space)
            g = torch.autograd.grad(loss, x)[0]
            prev = prev - rate * g
        return SimpleNamespace(prev_sample=prev.detach(), pred_original_sample=x0.detach(), loss=loss.detach())

    if kind == "mpgd":  # scheduling_mpgd.py:164-218
        x0 = base_x0(sample)
        with torch.enable_grad():
            x0 = x0.clone().detach().requires_grad_(True)
            loss = _residual_norm(operator, x0, vae, vocoder, L, measurement, supervised_space)
            g = torch.autograd.grad(loss, x0)[0]
            x0 = x0.detach() - rate * g
        e = (sample - a_t ** 0.5 * x0) / b_t ** 0.5
        prev = add_noise(a_prev ** 0.5 * x0 + (1 - a_prev - std ** 2) ** 0.5 * e)
        return SimpleNamespace(prev_sample=prev.detach(), pred_original_sample=x0, loss=loss.detach())

    if kind in ("dsg", "diffmusic"):  # scheduling_dsg.py:174-224 / scheduling_diffmusic.py:176-223
        with torch.enable_grad():
            x = sample.clone().detach().requires_grad_(True)
            x0 = base_x0(x, with_eta=False)
            mean = a_prev ** 0.5 * x0 + (1 - a_prev - std ** 2) ** 0.5 * model_output
            loss = _residual_norm(operator, x0, vae, vocoder, L, measurement, supervised_space)
            g = torch.autograd.grad(loss / 1000, x)[0]
            gn = torch.linalg.norm(g)
            if kind == "dsg":
                _, c, h, w = x.shape
                r = torch.sqrt(torch.tensor(c * h * w)) * std
                d_star = -r * g / (gn + eps)
                z = randn_like_reference(model_output.shape, generator, model_output.device, model_output.dtype)
                d_s = std * z
                mix = d_s + rate * (d_star - d_s)
                prev = mean + r * mix / (torch.linalg.norm(mix) + eps)
            else:
                z = randn_like_reference(model_output.shape, generator, model_output.device, model_output.dtype)
                gt = g / (gn + eps) * torch.linalg.norm(z)
                prev = mean + std * slerp(z, -gt, rate)
        return SimpleNamespace(prev_sample=prev.detach(), pred_original_sample=x0.detach(), loss=loss.detach())

    raise ValueError(f"Unknown scheduler: {kind}")


def per_clip_step(kind, base, operator, model_output, timestep, sample, generators=None, variance_noise=None,
                  measurement=None, **kw):
    """B independent batch-1 reference steps.  `generators`: list of B generators or None.
    `measurement` may be (1, ...) (shared) or (B, ...)."""
    B = sample.shape[0]
    prevs, x0s, losses = [], [], []
    for i in range(B):
        m = measurement
        if measurement is not None and measurement.shape[0] == B and B > 1:
            m = measurement[i:i + 1]
        out = reference_step(kind, base, operator, model_output[i:i + 1], timestep, sample[i:i + 1],
                             generator=None if generators is None else generators[i],
                             variance_noise=None if variance_noise is None else variance_noise[i:i + 1],
                             measurement=m, **kw)
        prevs.append(out.prev_sample)
        x0s.append(out.pred_original_sample)
        losses.append(out.loss.reshape(1).float())
    return SimpleNamespace(prev_sample=torch.cat(prevs), pred_original_sample=torch.cat(x0s),
                           loss=torch.cat(losses))
