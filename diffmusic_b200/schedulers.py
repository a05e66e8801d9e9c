"""Guided schedulers -- host-side mirror of diffmusic/schedulers/scheduling_{ddim,dps,mpgd,dsg,diffmusic}.py with the
same class names, constructor and `.step` signatures, running the latent algebra on the sm_100a kernels of
csrc/sched_update.cu and the measurement path on the fused guidance kernels (operators.py).

Batched semantics are per clip: a batch of B clips behaves like B independent batch-1 reference
trajectories (per-clip loss, per-clip gradient norms, per-clip slerp).  For B = 1 this is the reference.

The base class is diffusers' DDIMScheduler when diffusers is importable (so the reference pipelines accept the
object), otherwise the dependency-free restatement in ddim_base.py.  `super().step` is never called: x0 comes from
dm_sched_x0; the base's only other effect -- one discarded randn draw when eta > 0 -- is reproduced so generator
streams stay aligned with the reference.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib
from .ddim_base import (BaseOutput, DDIMBase, randn_clips_f32, randn_tensor, register_to_config, skip_randn)
from .operators import BaseOperator, generic_guidance_loss


@dataclass
class InverseProblemSchedulerOutput(BaseOutput):
    """diffmusic/schedulers/utils.py:8-16 (+ loss_per_clip for batched use)."""
    sample: Optional[torch.Tensor] = None
    prev_sample: torch.Tensor = None
    pred_original_sample: Optional[torch.Tensor] = None
    loss: Optional[torch.Tensor] = None
    encoder_hidden_states: Optional[torch.Tensor] = None
    encoder_hidden_states_1: Optional[torch.Tensor] = None
    init_latents: Optional[torch.Tensor] = None
    loss_per_clip: Optional[torch.Tensor] = None


def _f(t):
    """exact Python float of a 0-d fp32 tensor / number (so the kernel sees the reference's fp32 scalar)."""
    return float(t.item()) if isinstance(t, torch.Tensor) else float(t)


_LEAF_SCALE_CACHE = {}


class _GuidedBase(DDIMBase):
    """Shared constructor (scheduling_dps.py:22-61, identical in all five files) and helpers."""

    @register_to_config
    def __init__(self, operator: BaseOperator = None, num_train_timesteps: int = 1000, beta_start: float = 0.0001,
                 beta_end: float = 0.02, beta_schedule: str = "linear", trained_betas=None, clip_sample: bool = True,
                 set_alpha_to_one: bool = True, steps_offset: int = 0, prediction_type: str = "epsilon",
                 thresholding: bool = False, dynamic_thresholding_ratio: float = 0.995,
                 clip_sample_range: float = 1.0, sample_max_value: float = 1.0, timestep_spacing: str = "leading",
                 rescale_betas_zero_snr: bool = False, *args, **kwargs):
        super().__init__(num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
                         beta_schedule=beta_schedule, trained_betas=trained_betas, clip_sample=clip_sample,
                         set_alpha_to_one=set_alpha_to_one, steps_offset=steps_offset,
                         prediction_type=prediction_type, thresholding=thresholding,
                         dynamic_thresholding_ratio=dynamic_thresholding_ratio, clip_sample_range=clip_sample_range,
                         sample_max_value=sample_max_value, timestep_spacing=timestep_spacing,
                         rescale_betas_zero_snr=rescale_betas_zero_snr)
        self.operator = operator

    # ---- coefficient prep (scheduling_dps.py:157-162): 0-d fp32 CPU tensors, exactly the reference's arithmetic ----
    def _coeffs(self, timestep, eta):
        """per-timestep scalars, memoised: the ~15 0-d torch ops below cost more host time than a whole graph replay
        takes on the device, and the guided loop visits each (timestep, eta) once per clip batch."""
        key = (int(timestep), float(eta), self.num_inference_steps)
        cache = self.__dict__.setdefault("_coef_cache", {})
        if key not in cache:
            cache[key] = self._coeffs_uncached(timestep, eta)
        return cache[key]

    def _coeffs_uncached(self, timestep, eta):
        t = int(timestep)
        t_prev = t - self.config.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        b_t = 1 - a_t
        a_prev = self.alphas_cumprod[t_prev] if t_prev >= 0 else self.final_alpha_cumprod
        var = self._get_variance(t, t_prev)
        std = eta * var ** 0.5
        return dict(sqrt_a=_f(a_t ** 0.5), sqrt_b=_f(b_t ** 0.5), sqrt_p=_f(a_prev ** 0.5),
                    sqrt_1mp=_f((1 - a_prev) ** 0.5), dir_coef=_f((1 - a_prev - std ** 2) ** 0.5), std=_f(std))

    # ---- CUDA-graph support: when `_coef_dev` is set (GraphedGuidedStep), kernels read the per-timestep scalars from
    # this device tensor instead of the by-value arguments, so one captured graph serves every timestep ----
    _coef_dev = None

    def _coef_ptr(self):
        return None if self._coef_dev is None else self._coef_dev.data_ptr()

    def coef_vector(self, timestep, eta, n_clip):
        """[sqrt_a, sqrt_b, sqrt_p, dir_coef, std, r, 0, 0] as fp32 -- the layout dm_sched_* read from `coef`."""
        key = ("vec", int(timestep), float(eta), self.num_inference_steps, int(n_clip))
        cache = self.__dict__.setdefault("_coef_cache", {})
        if key not in cache:
            c = self._coeffs(timestep, eta)
            dirc = c["sqrt_1mp"] if isinstance(self, DDIMScheduler) else c["dir_coef"]
            r = _f(torch.sqrt(torch.tensor(n_clip)) * c["std"])
            cache[key] = torch.tensor([c["sqrt_a"], c["sqrt_b"], c["sqrt_p"], dirc, c["std"], r, 0.0, 0.0],
                                      dtype=torch.float32)
        return cache[key]

    def _check_supported(self):
        if self.config.prediction_type != "epsilon" or self.config.thresholding:
            raise NotImplementedError("the fused x0 kernel covers prediction_type='epsilon' without dynamic "
                                      "thresholding (every shipped config, configs/model/*.yaml)")

    @staticmethod
    def _prep_pair(sample, model_output):
        """(x, eps, io): the two latent-typed inputs as the kernels read them.  fp32 / fp16 / bf16 are consumed as they
        are (dm_sched_*_io convert on load and store); anything else, or mixed dtypes, goes through
        fp32."""
        _lib.require_cuda(sample, model_output)
        x, e = sample.detach(), model_output.detach()
        if x.dtype != e.dtype or x.dtype not in _lib.IO_DTYPES:
            x, e = x.float(), e.float()
        return x.contiguous(), e.contiguous(), _lib.IO_DTYPES[x.dtype]

    def _x0(self, x, eps, c, io, publish=True, leaf_scale=None):
        """dm_sched_x0_io: pred_original_sample of the diffusers base step.  Returns (x0 fp32 kept by the step,
        x0 in the latent dtype for the caller -- the same tensor for fp32 pipelines, leaf): `leaf` = leaf_scale * x0
        in the latent dtype (the VAE decoder input of scheduling_dps.py:195-197) when leaf_scale is given."""
        self._check_supported()
        x0 = torch.empty(x.shape, device=x.device, dtype=torch.float32)
        pub = torch.empty_like(x) if (io != 0 and publish) else None
        leaf = torch.empty_like(x) if leaf_scale is not None else None
        _lib.call("dm_sched_x0_io", x.data_ptr(), eps.data_ptr(), x0.data_ptr(), _lib.ptr(pub), _lib.ptr(leaf),
                  float(leaf_scale or 1.0), x.numel(), c["sqrt_a"], c["sqrt_b"], int(bool(self.config.clip_sample)),
                  float(self.config.clip_sample_range), self._coef_ptr(), io, _lib.stream())
        return x0, (x0 if pub is None else pub), leaf

    def _clip_grad_mask(self, g0, x, eps, c):
        """clip_sample=True (the constructor default; every shipped config sets False): DPS / DSG / DiffMusic
        differentiate w.r.t. `sample` THROUGH the base step's `pred_original_sample.clamp(+-clip_sample_range)`
        (scheduling_dps.py:165-175), so autograd zeroes the gradient where the unclipped x0 lies outside the range
        (clamp passes it at equality).  The fused update kernels start from dLoss/dx0 of the clipped x0, so the mask is
        applied to that gradient here; the unclipped x0 is recomputed with the arithmetic of dm_sched_x0.  Plain torch
        elementwise work on a configuration outside the shipped ones (MPGD leafs the clipped x0: no mask)."""
        if not self.config.clip_sample:
            return g0
        if self._coef_dev is not None:  # graph mode: per-timestep scalars live on the device
            sa, sb = self._coef_dev[0], self._coef_dev[1]
        else:
            sa, sb = c["sqrt_a"], c["sqrt_b"]
        t = (x.float() - sb * eps.float()) / sa
        r = float(self.config.clip_sample_range)
        keep = (t >= -r) & (t <= r)
        return (g0 * keep.to(g0.dtype)).contiguous()

    @staticmethod
    def _leaf_scale(vae):
        """1 / vae.config.scaling_factor as the fp32 value torch multiplies by (scheduling_dps.py:195-197)."""
        sf = vae.config.scaling_factor
        v = _LEAF_SCALE_CACHE.get(sf)
        if v is None:
            v = _LEAF_SCALE_CACHE[sf] = float(torch.tensor(1 / sf, dtype=torch.float32))
        return v

    #: which noise the step consumes: "eta" = base-step draw (discarded) + own z when eta > 0 (DDIM/DPS/MPGD);
    #: "always" = exactly one z per step, independent of eta (DSG/DiffMusic, scheduling_dsg.py:215-220)
    noise_mode = "eta"

    def draw_step_noise(self, eta, generator, variance_noise, model_output, out=None):
        """All RNG consumption of one step, in the reference's order, done up front (the values do not depend on
        where in the step they are drawn; only the per-generator order matters and is kept):
          DDIM/DPS/MPGD, eta > 0: the diffusers base step draws one tensor the reference discards
          (scheduling_dps.py:166-175), then the step's own z (:180-191); a given `variance_noise` replaces both draws.
          DSG/DiffMusic: one z per step (base step is called without eta, so it never draws).
        Returns z as fp32 (or None)."""
        shape, dev, dt = model_output.shape, model_output.device, model_output.dtype
        if self.noise_mode == "always":
            return self._randn_f32(shape, generator, dev, dt, out)
        if not eta > 0:
            return None
        if variance_noise is not None and generator is not None:
            raise ValueError("Cannot pass both generator and variance_noise. Please make sure that either "
                             "`generator` or `variance_noise` stays `None`.")
        if variance_noise is None:
            skip_randn(shape, generator, device=dev, dtype=dt)  # base-step draw, discarded by the reference
            if isinstance(self, DDIMScheduler):
                return None
            return self._randn_f32(shape, generator, dev, dt, out)
        return variance_noise.detach().float().contiguous()

    @staticmethod
    def _randn_f32(shape, generator, dev, dt, out=None):
        """the step noise as fp32 (written into `out` when given and possible): per-clip generator lists in one launch
        (ddim_base.randn_clips_f32), else torch"""
        z = randn_clips_f32(shape, generator, dev, dt, out=out)
        if z is None:
            z = randn_tensor(shape, generator=generator, device=dev, dtype=dt).float().contiguous()
        return z

    def _noise_arg(self, given, eta, generator, variance_noise, model_output):
        """step noise: drawn here (eager) unless the caller already did (`_noise`, used by GraphedGuidedStep)."""
        if given is None:
            return self.draw_step_noise(eta, generator, variance_noise, model_output)
        return given if isinstance(given, torch.Tensor) else None

    def _guidance(self, leaf, measurement, vae, vocoder, L, supervised_space, model_dtype=None):
        """loss (per clip) and dLoss/d(leaf) through vae.decode + vocoder (torch autograd) and the fused operator
        kernels (scheduling_dps.py:195-212).  `leaf` = 1/scaling_factor * x0 from dm_sched_x0_io, in the pipeline's
        dtype: autograd starts there (the UNet is never differentiated; the scaling and its backward are
        inside our kernels), and the gradient goes to the update kernel as autograd returns it."""
        if supervised_space not in ("wav_form", "mel_spectrogram"):
            raise ValueError("supervised_space should be either 'wav_form' or 'mel_spectrogram")
        op = self.operator
        io_dtype = leaf.dtype
        with torch.enable_grad():
            leaf = leaf.detach()
            if model_dtype is not None and leaf.dtype != model_dtype:  # exotic / mixed input dtypes went through fp32
                leaf = leaf.to(model_dtype)
            leaf.requires_grad_(True)
            mel = vae.decode(leaf).sample
            wav_full = op.inverse_transform(mel, vocoder)
            wav = wav_full[:, :L]
            if getattr(op, "fused_loss_and_grad", None) is not None and wav_full.dim() == 2 \
                    and wav_full.stride(1) == 1 and (wav_full.dtype == torch.float32 or op.wave16_ok(wav)):
                # one fused kernel chain gives the per-clip loss AND dLoss/dwav; torch only continues the VJP through
                # the vocoder and the VAE decoder.  The cotangent is written straight into a buffer shaped like the
                # vocoder output (zero tail beyond L = the adjoint of the `[:, :L]` slice), so autograd
                # starts at the vocoder output: no slice-backward zero-fill + copy of the whole waveform.  fp16 / bf16
                # pipelines: the kernels read the 16-bit waveform and write the 16-bit cotangent directly.
                Lw = wav.shape[1]
                dfull = self._cotangent_buffer(wav_full, Lw)
                losses, _ = op.fused_loss_and_grad(wav.detach(), measurement, supervised_space, dwav=dfull[:, :Lw])
                (g,) = torch.autograd.grad(wav_full, leaf, grad_outputs=dfull)
            elif getattr(op, "fused_loss_and_grad", None) is not None:
                losses, dwav = op.fused_loss_and_grad(wav.detach(), measurement, supervised_space)
                (g,) = torch.autograd.grad(wav, leaf, grad_outputs=dwav.to(wav.dtype))
            else:
                if hasattr(op, "guidance_loss"):
                    losses = op.guidance_loss(wav, measurement, supervised_space)
                else:
                    losses = generic_guidance_loss(op, wav, measurement, supervised_space)
                (g,) = torch.autograd.grad(losses.sum(), leaf)
        return losses.detach(), g.to(io_dtype).contiguous()

    def _cotangent_buffer(self, wav_full, Lw):
        """vocoder-shaped cotangent buffer, kept between steps: the kernels overwrite [:, :Lw] every step and never touch
        the tail, which is zeroed once here (no per-step fill kernel).  A buffer that a CUDA graph captured is never
        released (the graph replays into its address); otherwise at most a handful of shapes are kept."""
        key = (tuple(wav_full.shape), wav_full.dtype, wav_full.device, Lw)
        cache = self.__dict__.setdefault("_dfull_cache", {})
        captured = self.__dict__.setdefault("_dfull_captured", set())
        buf = cache.get(key)
        if buf is None:
            for old in [k for k in cache if k not in captured][:max(0, len(cache) - 3)]:
                del cache[old]
            buf = torch.zeros_like(wav_full)
            cache[key] = buf
        if torch.cuda.is_current_stream_capturing():
            captured.add(key)
        return buf

    @staticmethod
    def _loss_slot(losses):
        """(contiguous fp32 per-clip losses, 0-d device tensor the update kernel fills with their 2-norm): the
        reference's 0-d `loss` = torch.linalg.norm over the whole batch (= the per-clip norm for B = 1)."""
        losses = losses.float().contiguous()
        return losses, torch.empty((), device=losses.device, dtype=torch.float32)

    def optim_prompt(self, model_output, timestep, sample, encoder_hidden_states=None, encoder_hidden_states_1=None,
                     eta=0.0, use_clipped_model_output=False, generator=None, variance_noise=None, return_dict=True,
                     measurement=None, vae=None, vocoder=None, original_waveform_length=0,
                     optim_prompt_learning_rate=1e-4, supervised_space="mel_spectrogram", *args, **kwargs):
        """scheduling_dps.py:63-135.  Disabled in every shipped config (configs/*.yaml: optim_prompt: false) and, as
        written in the reference, a no-op on the embeddings (the SGD steps act on discarded clones, :94-96): it returns
        the embeddings detached.  Kept for API compatibility; the guided loss is still evaluated so errors surface."""
        c = self._coeffs(timestep, eta)
        x, eps, io = self._prep_pair(sample, model_output)
        _, _, leaf = self._x0(x, eps, c, io, publish=False, leaf_scale=self._leaf_scale(vae))
        self._guidance(leaf, measurement, vae, vocoder, original_waveform_length, supervised_space, sample.dtype)
        return InverseProblemSchedulerOutput(
            encoder_hidden_states=None if encoder_hidden_states is None else encoder_hidden_states.detach(),
            encoder_hidden_states_1=None if encoder_hidden_states_1 is None else encoder_hidden_states_1.detach())


class DDIMScheduler(_GuidedBase):
    """scheduling_ddim.py:58-104.  Unlike the reference (which raises when the pipelines call it without
    encoder_hidden_states) None is accepted for the two passthrough tensors."""

    def step(self, model_output, timestep, sample, eta: float = 0.0, use_clipped_model_output: bool = False,
             generator=None, variance_noise=None, return_dict: bool = True, measurement=None, vae=None, vocoder=None,
             original_waveform_length: int = 0, encoder_hidden_states=None, encoder_hidden_states_1=None, *args,
             _noise=None, **kwargs):
        c = self._coeffs(timestep, eta)
        x, eps, io = self._prep_pair(sample, model_output)
        self._noise_arg(_noise, eta, generator, variance_noise, model_output)
        x0, x0_pub, _ = self._x0(x, eps, c, io)
        prev = torch.empty_like(x)
        _lib.call("dm_sched_ddim_update_io", x.data_ptr(), x0.data_ptr(), prev.data_ptr(), x.numel(), c["sqrt_a"],
                  c["sqrt_b"], c["sqrt_p"], c["sqrt_1mp"], self._coef_ptr(), io, _lib.stream())
        return InverseProblemSchedulerOutput(
            prev_sample=prev.to(sample.dtype), pred_original_sample=x0_pub.to(sample.dtype),
            loss=torch.tensor([int(timestep)]),
            encoder_hidden_states=None if encoder_hidden_states is None else encoder_hidden_states.detach(),
            encoder_hidden_states_1=None if encoder_hidden_states_1 is None else encoder_hidden_states_1.detach())


class DPSScheduler(_GuidedBase):
    """Diffusion Posterior Sampling, scheduling_dps.py:137-219."""

    def step(self, model_output, timestep, sample, eta: float = 0.0, use_clipped_model_output: bool = False,
             generator=None, variance_noise=None, return_dict: bool = True, measurement=None,
             ip_guidance_rate: float = 5e-4, vae=None, vocoder=None, original_waveform_length: int = 0,
             supervised_space: str = "mel_spectrogram", *args, _noise=None, **kwargs):
        c = self._coeffs(timestep, eta)
        x, eps, io = self._prep_pair(sample, model_output)
        z = self._noise_arg(_noise, eta, generator, variance_noise, model_output)
        ls = self._leaf_scale(vae)
        x0, x0_pub, leaf = self._x0(x, eps, c, io, leaf_scale=ls)
        losses, g0 = self._guidance(leaf, measurement, vae, vocoder, original_waveform_length, supervised_space,
                                    sample.dtype)
        g0 = self._clip_grad_mask(g0, x, eps, c)
        prev = torch.empty_like(x)
        losses, total = self._loss_slot(losses)
        _lib.call("dm_sched_dps_update_io", x.data_ptr(), x0.data_ptr(), g0.data_ptr(), ls, _lib.ptr(z),
                  prev.data_ptr(), x.numel(), c["sqrt_a"], c["sqrt_b"], c["sqrt_p"], c["dir_coef"], c["std"],
                  float(ip_guidance_rate), self._coef_ptr(), io, losses.data_ptr(), losses.numel(), total.data_ptr(),
                  _lib.stream())
        return InverseProblemSchedulerOutput(prev_sample=prev.to(sample.dtype),
                                             pred_original_sample=x0_pub.to(sample.dtype),
                                             loss=total, loss_per_clip=losses)


class MPGDScheduler(_GuidedBase):
    """Manifold Preserving Guided Diffusion, scheduling_mpgd.py:137-224."""

    def step(self, model_output, timestep, sample, eta: float = 0.0, use_clipped_model_output: bool = False,
             generator=None, variance_noise=None, return_dict: bool = True, measurement=None,
             ip_guidance_rate: float = 1.0, vae=None, vocoder=None, original_waveform_length: int = 0,
             supervised_space: str = "mel_spectrogram", *args, _noise=None, **kwargs):
        c = self._coeffs(timestep, eta)
        x, eps, io = self._prep_pair(sample, model_output)
        z = self._noise_arg(_noise, eta, generator, variance_noise, model_output)
        ls = self._leaf_scale(vae)
        x0, _, leaf = self._x0(x, eps, c, io, publish=False, leaf_scale=ls)
        losses, g0 = self._guidance(leaf, measurement, vae, vocoder, original_waveform_length, supervised_space,
                                    sample.dtype)
        prev, x0_new = torch.empty_like(x), torch.empty_like(x)
        losses, total = self._loss_slot(losses)
        _lib.call("dm_sched_mpgd_update_io", x.data_ptr(), x0.data_ptr(), g0.data_ptr(), ls, _lib.ptr(z),
                  prev.data_ptr(), x0_new.data_ptr(), x.numel(), c["sqrt_a"], c["sqrt_b"], c["sqrt_p"], c["dir_coef"],
                  c["std"], float(ip_guidance_rate), self._coef_ptr(), io, losses.data_ptr(), losses.numel(),
                  total.data_ptr(), _lib.stream())
        return InverseProblemSchedulerOutput(prev_sample=prev.to(sample.dtype),
                                             pred_original_sample=x0_new.to(sample.dtype),
                                             loss=total, loss_per_clip=losses)


class _SphericalBase(_GuidedBase):
    """DSG and DiffMusic share everything up to the per-clip mixing rule."""

    noise_mode = "always"

    def _spherical_step(self, kernel, model_output, timestep, sample, eta, generator, variance_noise, measurement,
                        vae, vocoder, L, ip_guidance_rate, eps, supervised_space, _noise=None):
        c = self._coeffs(timestep, eta)
        x, e, io = self._prep_pair(sample, model_output)
        # one draw per step (scheduling_dsg.py:215-220; `variance_noise` is not consulted by the reference there)
        z = self._noise_arg(_noise, eta, generator, variance_noise, model_output)
        ls = self._leaf_scale(vae)
        # base step called without eta (scheduling_dsg.py:178-186): no RNG draw
        x0, x0_pub, leaf = self._x0(x, e, c, io, leaf_scale=ls)
        losses, g0 = self._guidance(leaf, measurement, vae, vocoder, L, supervised_space, sample.dtype)
        g0 = self._clip_grad_mask(g0, x, e, c)
        B = x.shape[0]
        n_clip = x.numel() // B
        prev = torch.empty_like(x)
        losses, total = self._loss_slot(losses)
        if kernel == "dsg":
            # r = sqrt(c*h*w) * std as fp32 (scheduling_dsg.py:212-213)
            r = _f(torch.sqrt(torch.tensor(n_clip)) * c["std"])
            _lib.call("dm_sched_dsg_update_io", x0.data_ptr(), e.data_ptr(), g0.data_ptr(), ls, z.data_ptr(),
                      prev.data_ptr(), B, n_clip, c["sqrt_a"], c["sqrt_p"], c["dir_coef"], c["std"],
                      float(ip_guidance_rate), r, 1.0 / 1000.0, float(eps), self._coef_ptr(), io, losses.data_ptr(),
                      total.data_ptr(), _lib.stream())
        else:
            _lib.call("dm_sched_diffmusic_update_io", x0.data_ptr(), e.data_ptr(), g0.data_ptr(), ls, z.data_ptr(),
                      prev.data_ptr(), B, n_clip, c["sqrt_a"], c["sqrt_p"], c["dir_coef"], c["std"],
                      float(ip_guidance_rate), 1.0 / 1000.0, float(eps), 0.9995, self._coef_ptr(), io,
                      losses.data_ptr(), total.data_ptr(), _lib.stream())
        return InverseProblemSchedulerOutput(prev_sample=prev.to(sample.dtype),
                                             pred_original_sample=x0_pub.to(sample.dtype),
                                             loss=total, loss_per_clip=losses)


class DSGScheduler(_SphericalBase):
    """Diffusion with Spherical Gaussian constraint, scheduling_dsg.py:148-230.  eta defaults to 1."""

    def step(self, model_output, timestep, sample, eta: float = 1.0, use_clipped_model_output: bool = False,
             generator=None, variance_noise=None, return_dict: bool = True, measurement=None, vae=None, vocoder=None,
             original_waveform_length: int = 0, ip_guidance_rate: float = 0.08, eps: float = 1e-8,
             supervised_space: str = "mel_spectrogram", *args, _noise=None, **kwargs):
        return self._spherical_step("dsg", model_output, timestep, sample, eta, generator, variance_noise,
                                    measurement, vae, vocoder, original_waveform_length, ip_guidance_rate, eps,
                                    supervised_space, _noise)


class DiffMusicScheduler(_SphericalBase):
    """scheduling_diffmusic.py:148-229 + slerp :59-68; the |cos| > 0.9995 branch is taken on the
    device per clip, so `.step` never synchronises with the host."""

    @staticmethod
    def slerp(x0, x1, gamma=0.008, threshold=0.9995):
        """scheduling_diffmusic.py:59-68 (public static helper of the reference class; torch, for API parity)."""
        cos_theta = ((x0 / torch.norm(x0)) * (x1 / torch.norm(x1))).sum()
        if cos_theta.abs() > threshold:
            return x0 + gamma * (x1 - x0)
        theta = torch.acos(cos_theta)
        sin_theta = torch.sin(theta)
        return torch.sin((1 - gamma) * theta) / sin_theta * x0 + torch.sin(gamma * theta) / sin_theta * x1

    def step(self, model_output, timestep, sample, eta: float = 0.0, use_clipped_model_output: bool = False,
             generator=None, variance_noise=None, return_dict: bool = True, measurement=None, vae=None, vocoder=None,
             original_waveform_length: int = 0, ip_guidance_rate: float = 0.08, eps: float = 1e-8,
             supervised_space: str = "mel_spectrogram", *args, _noise=None, **kwargs):
        return self._spherical_step("diffmusic", model_output, timestep, sample, eta, generator, variance_noise,
                                    measurement, vae, vocoder, original_waveform_length, ip_guidance_rate, eps,
                                    supervised_space, _noise)


def _reference_ditto():
    """The reference's own DITTOScheduler.  With the drop-in ahead of the reference on sys.path the package
    `diffmusic.schedulers` is ours (its search path is extended with the reference's directory, dropin/diffmusic/
    schedulers/__init__.py); should the import still not resolve, the file is located on sys.path and loaded directly."""
    import importlib
    import importlib.util
    import os
    import sys
    try:
        return importlib.import_module("diffmusic.schedulers.scheduling_ditto").DITTOScheduler
    except ModuleNotFoundError as exc:
        if exc.name not in ("diffmusic", "diffmusic.schedulers", "diffmusic.schedulers.scheduling_ditto"):
            raise  # the reference file was found, one of ITS dependencies (e.g. diffusers) is missing
        first = exc
    for entry in sys.path:
        path = os.path.join(entry or ".", "diffmusic", "schedulers", "scheduling_ditto.py")
        if os.path.isfile(path):
            spec = importlib.util.spec_from_file_location("diffmusic.schedulers.scheduling_ditto", path)
            mod = importlib.util.module_from_spec(spec)
            sys.modules[spec.name] = mod
            try:
                spec.loader.exec_module(mod)
            except BaseException:
                sys.modules.pop(spec.name, None)
                raise
            return mod.DITTOScheduler
    raise ImportError("DITTOScheduler is not part of diffmusic_b200; put the reference's "
                      "diffmusic/schedulers/scheduling_ditto.py on sys.path to use it") from first


def get_scheduler(scheduler_name):
    """diffmusic/schedulers/__init__.py:9-24."""
    table = {"ddim": DDIMScheduler, "dps": DPSScheduler, "mpgd": MPGDScheduler, "dsg": DSGScheduler,
             "diffmusic": DiffMusicScheduler}
    if scheduler_name in table:
        return table[scheduler_name]
    if scheduler_name == "ditto":
        # DITTO back-propagates through the whole sampling chain including the UNet (scheduling_ditto.py:187-208):
        # out of the hot-path scope, so the reference class is re-exported (diffmusic/schedulers/__init__.py:19-20).
        return _reference_ditto()
    raise ValueError(f"Unknown scheduler: {scheduler_name}")
