"""Measurement operators A(x), their transforms and the fused guidance loss -- host-side mirror of
diffmusic/inverse_problem/operator.py (same class names, constructor arguments and methods), running on the
hand-written sm_100a kernels behind include/dm_abi.h.

  forward(data)            -> A(x)                      (operator.py:44-45,132-133,162-171,203-205,244-250)
  transform(x)             -> wav -> dB-mel / mag -> mel (operator.py:35-36,123-124,153-154,194-195,229-230)
  inverse_transform(m, v)  -> vocoder(mel)              (operator.py:38-42; the vocoder stays in PyTorch)
  guidance_loss(wav, measurement, supervised_space) -> per-clip ||residual||_2 with a fused VJP (new; what the
                              schedulers call instead of forward + transform + sub + linalg.norm + autograd replay,
                              scheduling_dps.py:198-212)

There is no CPU implementation here.  CPU tensors given to forward/transform (run.py:286,312 builds the measurement
on the CPU) are moved to the GPU, processed by the same kernels and moved back.
"""
from __future__ import annotations

import ctypes as C
import math
import os

import torch

from . import _lib, tables

_MODE_MEL_DB, _MODE_PHASE_MEL, _MODE_PHASE_WAV = 0, 1, 2
_RESID_CHUNK = 4096
_RIR_MAX_TAPS = 6144
_HOP = 160


def _device_of(t):
    if t.is_cuda:
        return t.device
    if not torch.cuda.is_available():
        raise _lib.DiffMusicB200Error("diffmusic_b200 needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def _as_f32_rows(t):
    """fp32 2-D view with unit inner stride (row stride free); copies only when it has to."""
    if t.dtype != torch.float32:
        t = t.float()
    if t.dim() == 1:
        t = t.unsqueeze(0)
    if t.stride(-1) != 1:
        t = t.contiguous()
    return t


def _wave_rows(t):
    """2-D rows with unit inner stride in a dtype the waveform-typed kernels read directly (fp32 / fp16 / bf16);
    returns (rows, DM_IO code).  Other dtypes go through fp32."""
    if t.dtype not in _lib.IO_DTYPES:
        t = t.float()
    if t.dim() == 1:
        t = t.unsqueeze(0)
    if t.stride(-1) != 1:
        t = t.contiguous()
    return t, _lib.IO_DTYPES[t.dtype]


class _DeviceTables:
    """Per-device constant tables (window, twiddles, sparse mel tables) and the ctypes struct pointing at them."""

    def __init__(self, device, window, sample_rate=16000):
        fb = tables.mel_filterbank(sample_rate)
        mt = tables.mel_tables(fb)
        self.t = {"window": window.to(device), "tw512": tables.twiddles(512).to(device),
                  "w1024": tables.half_twiddles(1024).to(device)}
        for k, v in mt.items():
            self.t[k] = v.to(device).contiguous()
        img, na, nb = tables.warp_image(window, fb)
        self.t["warp_image"] = img.to(device).contiguous()
        p = {k: v.data_ptr() for k, v in self.t.items()}
        self.struct = _lib.StftTables(p["window"], p["tw512"], p["w1024"], p["mel_kstart"], p["mel_klen"], p["mel_w"],
                                      tables.MEL_WSTRIDE, p["bin_m0"], p["bin_w0"], p["bin_w1"], p["warp_image"],
                                      img.numel(), na, nb)
        self.ref = C.byref(self.struct)


def _grad_buffer(dwav, B, L, device, dtype=torch.float32):
    """(B, L) destination of dLoss/dwav in the waveform's dtype: a fresh tensor, or the caller's (possibly row-strided)
    view."""
    if dwav is None:
        return torch.empty((B, L), device=device, dtype=dtype)
    if tuple(dwav.shape) != (B, L) or dwav.dtype != dtype or dwav.stride(1) != 1 or dwav.device != device:
        raise ValueError(f"dwav must be a (B, L) {dtype} view with unit inner stride on the waveform's device")
    return dwav


def _frames_per_tile(B, T, device, max_nf=16):
    """frames per CTA tile of the STFT guidance kernel (`max_nf` = 14 for the fused resampling chain, whose last warp
    owns no frame pair).

    The warp-per-frame-pair engine (csrc/stft_warp.cu) gives each of its 8 warps one frame pair, so a tile holds at most
    16 frames; its persistent grid is 2 CTAs per SM.  Cost model: a CTA walks `ceil(ctas / slots)` tiles, a tile costs
    its frames plus a fixed staging / barrier overhead; pick the cheapest even size, larger tiles on ties (fewer
    re-staged samples and overlapping cotangent adds).  6 frames is the minimum for the bit-reproducible two-tile overlap
    of the cotangent.  `DM_STFT_FRAMES_PER_TILE` overrides the choice (tuning / tests; > 16 selects the 64-thread
    frame-pair kernel of stft_guidance.cu)."""
    forced = os.environ.get("DM_STFT_FRAMES_PER_TILE")
    if forced:
        return max(6, min(22 if max_nf == 16 else max_nf, int(forced)))
    key = (B, T, str(device), max_nf)
    if key in _NF_CACHE:
        return _NF_CACHE[key]
    slots = 2 * torch.cuda.get_device_properties(device).multi_processor_count
    best, best_cost = max_nf, float("inf")
    for nf in range(8, max_nf + 1, 2):
        ctas = B * math.ceil(T / nf)
        cost = math.ceil(ctas / slots) * (nf + 3)
        if cost <= best_cost + 1e-9:
            best, best_cost = nf, min(cost, best_cost)
    _NF_CACHE[key] = best
    return best


_NF_CACHE = {}


class BaseOperator:
    """operator.py:6-14 plus the shared kernel plumbing."""

    sample_rate = 16000
    #: the fused loss + VJP chain reads fp16 / bf16 waveforms and writes dLoss/dwav in that dtype directly
    wave16 = False
    clamp_transform = True     # every T_mel clamps to +-80 except MusicInpaintingOperator (operator.py:123-124)
    window_kind = "hann"
    noiser = None

    # ---- reference API -------------------------------------------------------------------------------------------
    def transform(self, data, *args, **kwargs):
        raise NotImplementedError

    def inverse_transform(self, mel_spectrogram, vocoder):
        """operator.py:38-42 (and its four clones)."""
        if mel_spectrogram.dim() == 4:
            mel_spectrogram = mel_spectrogram.squeeze(1)
        return vocoder(mel_spectrogram)

    def forward(self, data, *args, **kwargs):
        raise NotImplementedError

    # ---- plumbing ------------------------------------------------------------------------------------------------
    def _tables(self, device):
        cache = self.__dict__.setdefault("_tab_cache", {})
        key = str(device)
        if key not in cache:
            win = tables.hann_window() if self.window_kind == "hann" else tables.rect_window()
            cache[key] = _DeviceTables(device, win, self.sample_rate)
        return cache[key]

    def _sigma(self):
        n = self.noiser
        return float(getattr(n, "sigma", 0.0) or 0.0) if n is not None else 0.0

    def _stft(self, mode, y, *, mask=None, ref=None, out_rows=None, want_grad=False, clamp=None, noise=None,
              sigma=0.0, ypbar=None):
        """Launch dm_stft_guidance.  y: (B, Ly) fp32 rows on the GPU.
        transform mode (ref None) -> returns out (B, R, T); guidance mode -> returns (ypbar or None, partial, ntiles)."""
        B, Ly = y.shape
        dev = y.device
        y_io = _lib.IO_DTYPES[y.dtype]
        T = 1 + Ly // _HOP
        nf = _frames_per_tile(B, T, dev)
        ntiles = math.ceil(T / nf)
        tab = self._tables(dev)
        clamp = self.clamp_transform if clamp is None else clamp
        if ref is None:
            out = torch.empty((B, out_rows, T), device=dev, dtype=torch.float32)
            _lib.call("dm_stft_guidance_io", tab.ref, mode, int(clamp), _HOP, y.data_ptr(), y_io, y.stride(0), Ly,
                      _lib.ptr(mask), B, None, 0, _lib.ptr(noise), float(sigma), out.data_ptr(), None, None, nf,
                      _lib.stream())
            return out
        rows = 513 if mode == _MODE_PHASE_WAV else 64
        if tuple(ref.shape[1:]) != (rows, T) or ref.shape[0] not in (1, B):
            raise ValueError(f"measurement transform has shape {tuple(ref.shape)}, prediction needs (1|{B}, {rows}, {T})")
        ref_b = 0 if ref.shape[0] == 1 else ref.stride(0)
        partial = torch.empty((B, ntiles), device=dev, dtype=torch.float32)
        if not want_grad:
            ypbar = None
        elif ypbar is None:  # the kernel accumulates into it; a caller-provided buffer has been zeroed on the way
            ypbar = torch.zeros((B, Ly + 1024), device=dev, dtype=torch.float32)
        _lib.call("dm_stft_guidance_io", tab.ref, mode, int(clamp), _HOP, y.data_ptr(), y_io, y.stride(0), Ly,
                  _lib.ptr(mask), B, ref.data_ptr(), ref_b, _lib.ptr(noise), float(sigma), None, _lib.ptr(ypbar),
                  partial.data_ptr(), nf, _lib.stream())
        return ypbar, partial, ntiles

    def _mel_db(self, wav):
        """T_mel on (..., L) -> (..., 64, T)."""
        dev = _device_of(wav)
        lead = wav.shape[:-1]
        y = _as_f32_rows(wav.detach().to(dev).reshape(-1, wav.shape[-1]))
        out = self._stft(_MODE_MEL_DB, y, out_rows=64)
        return out.reshape(*lead, 64, out.shape[-1]).to(wav.device)

    def _finish_forward(self, y, like):
        """noiser(y) (noise.py:13-18) with the noise drawn by torch exactly where the reference draws it, then
        back to the caller's device."""
        sigma = self._sigma()
        if self.noiser is not None and not hasattr(self.noiser, "sigma"):
            return self.noiser(y.to(like.device))  # e.g. PoissonNoise: host numpy path of the reference, out of scope
        if self.noiser is not None:
            if not like.is_cuda:
                # the reference draws on the CPU for CPU inputs, even when sigma == 0 (keeps the CPU RNG stream equal)
                noise = torch.randn_like(y, device="cpu").to(y.device)
            elif sigma != 0.0:
                noise = torch.randn_like(y)
            else:
                noise = None
            if noise is not None and sigma != 0.0:
                _lib.call("dm_add_scaled", y.data_ptr(), noise.data_ptr(), sigma, y.numel(), _lib.stream())
        return y.to(like.device)

    # ---- fused guidance: per-clip loss with VJP --------------------------------------------------------------------
    def guidance_loss(self, wav, measurement, supervised_space="mel_spectrogram"):
        """per-clip ||measurement-space residual||_2, shape (B,), differentiable w.r.t. `wav` through one fused
        forward+VJP kernel chain (scheduling_dps.py:200-212).  measurement: (1, ...) shared or (B, ...)."""
        if supervised_space not in ("wav_form", "mel_spectrogram"):
            raise ValueError("supervised_space should be either 'wav_form' or 'mel_spectrogram")
        return _GuidanceLoss.apply(wav, self, measurement, supervised_space)

    def fused_loss_and_grad(self, wav, measurement, supervised_space="mel_spectrogram", dwav=None):
        """(loss (B,), dLoss/dwav (B, L)) from one fused forward+VJP kernel chain; what the schedulers call.
        `dwav` may be a preallocated (B, L) view with a row stride (e.g. the head of a (B, L_vocoder) buffer, so the
        gradient of the `[:, :L]` slice needs no extra zero-fill + copy)."""
        if supervised_space not in ("wav_form", "mel_spectrogram"):
            raise ValueError("supervised_space should be either 'wav_form' or 'mel_spectrogram")
        _lib.require_cuda(wav)
        rows = _wave_rows(wav)[0] if self.wave16_ok(wav) else _as_f32_rows(wav)
        return self._fused(rows, measurement, supervised_space, True, dwav)

    def wave16_ok(self, wav):
        """True when the fused chain can consume `wav` in its own 16-bit dtype (and will return dLoss/dwav in it)."""
        return (self.wave16 and wav.dtype in (torch.float16, torch.bfloat16) and wav.dim() == 2 and wav.stride(1) == 1
                and self._sigma() == 0.0)

    def _ref_mel(self, measurement):
        """transform(measurement), cached: the reference recomputes it every step (scheduling_dps.py:205)."""
        key = (measurement.data_ptr(), tuple(measurement.shape), measurement._version, str(measurement.device))
        cache = self.__dict__.setdefault("_ref_cache", {})
        if cache.get("key") != key:
            cache["key"] = key
            cache["val"] = self.transform(measurement.detach()).float().contiguous()
            cache["keep"] = measurement  # keep the storage alive so data_ptr cannot be recycled
        return cache["val"]

    # subclasses implement: _fused(wav_rows, measurement, space, want_grad) -> (loss (B,), dwav (B, L) or None)
    def _fused(self, wav, measurement, space, want_grad, dwav=None):
        raise NotImplementedError

    def _residual_wav(self, y, meas, mask=None):
        B, n = y.shape
        nt = math.ceil(n / _RESID_CHUNK)
        ybar = torch.empty((B, n), device=y.device, dtype=torch.float32)
        partial = torch.empty((B, nt), device=y.device, dtype=torch.float32)
        meas = _as_f32_rows(meas.reshape(meas.shape[0], -1))
        if meas.shape[1] != n:
            raise ValueError(f"measurement has {meas.shape[1]} samples per clip, prediction has {n}")
        _lib.call("dm_residual_wav_io", y.data_ptr(), _lib.IO_DTYPES[y.dtype], y.stride(0), n, B, _lib.ptr(mask),
                  meas.data_ptr(), 0 if meas.shape[0] == 1 else meas.stride(0), ybar.data_ptr(), partial.data_ptr(),
                  _lib.stream())
        return ybar, partial, nt

    def _fold_adjoint(self, ybar, pad, Ly, B, partial, ntiles, mask, want_grad, dwav=None, dtype=torch.float32):
        dev = partial.device
        loss = torch.empty((B,), device=dev, dtype=torch.float32)
        if not want_grad:  # loss only
            _lib.call("dm_fold_adjoint", None, pad, Ly, B, None, partial.data_ptr(), ntiles, None, 0,
                      loss.data_ptr(), _lib.stream())
            return loss, None
        dwav = _grad_buffer(dwav, B, Ly, dev, dtype)
        _lib.call("dm_fold_adjoint_io", ybar.data_ptr(), pad, Ly, B, _lib.ptr(mask), partial.data_ptr(), ntiles,
                  dwav.data_ptr(), _lib.IO_DTYPES[dtype], dwav.stride(0), loss.data_ptr(), _lib.stream())
        return loss, dwav

    def _space_stage(self, y, measurement, space, want_grad, mask=None, ypbar=None):
        """stage B: residual in `space` on y = A(x).  Returns (ybar, pad, partial, ntiles).  `ypbar`: an already
        zeroed (B, Ly + 1024) cotangent buffer for the mel space (else allocated and zeroed here)."""
        if space == "mel_spectrogram":
            ref = self._ref_mel(measurement).to(y.device)
            ypbar, partial, nt = self._stft(_MODE_MEL_DB, y, mask=mask, ref=ref, want_grad=want_grad, ypbar=ypbar)
            return ypbar, 512, partial, nt
        ybar, partial, nt = self._residual_wav(y, measurement.to(y.device), mask=mask)
        return ybar, 0, partial, nt


class _GuidanceLoss(torch.autograd.Function):
    """loss_b = ||residual_b||_2 ; backward = grad_out_b * dLoss_b/dwav, where dLoss/dwav was produced by the same
    kernel chain as the loss (the spectrum never leaves the SM, nothing is saved for backward but dwav)."""

    @staticmethod
    def forward(ctx, wav, op, measurement, space):
        _lib.require_cuda(wav)
        want_grad = ctx.needs_input_grad[0]
        rows = _as_f32_rows(wav.detach())
        loss, dwav = op._fused(rows, measurement, space, want_grad)
        ctx.dwav = dwav
        ctx.in_dtype = wav.dtype
        ctx.in_shape = wav.shape
        return loss

    @staticmethod
    def backward(ctx, grad_loss):
        if ctx.dwav is None:
            return None, None, None, None
        g = ctx.dwav * grad_loss.reshape(-1, 1).to(ctx.dwav.dtype)
        return g.reshape(ctx.in_shape).to(ctx.in_dtype), None, None, None


# ======================================================================================================== operators
class IdentityOperator(BaseOperator):
    """operator.py:17-45: forward = identity, transform = clamp(T_mel)."""

    def __init__(self, sample_rate):
        self.sample_rate = sample_rate

    def transform(self, audio):
        return self._mel_db(audio)

    def forward(self, data, **kwargs):
        return data

    wave16 = True

    def _fused(self, wav, measurement, space, want_grad, dwav=None):
        B, L = wav.shape
        ybar, pad, partial, nt = self._space_stage(wav, measurement, space, want_grad)
        return self._fold_adjoint(ybar, pad, L, B, partial, nt, None, want_grad, dwav, wav.dtype)


class MusicInpaintingOperator(BaseOperator):
    """operator.py:48-133.  The mask is built with the reference's integer index arithmetic (bit-exact); A(x) = x*mask
    is fused into the frame load of the STFT kernel and into the adjoint."""

    clamp_transform = False  # operator.py:123-124: no clamp for inpainting
    wave16 = True

    def __init__(self, audio_length_in_s, sample_rate, mask_type, start_inpainting_s, end_inpainting_s,
                 mask_percentage, mask_duration_s, interval_s, noiser=None):
        self.audio_length_in_s = audio_length_in_s
        self.sample_rate = sample_rate
        self.mask_type = mask_type
        self.start_inpainting_s = start_inpainting_s
        self.end_inpainting_s = end_inpainting_s
        self.mask_percentage = mask_percentage
        self.interval_s = interval_s
        self.mask_duration_s = mask_duration_s
        self.mask = self.generate_mask()
        self.noiser = noiser

    def generate_mask(self):
        """operator.py:87-121 -- (1, L) fp32 ones with zeros on the masked index ranges; `random` draws its window
        starts from the global CPU generator with torch.randint exactly like the reference."""
        sr = self.sample_rate
        n = self.audio_length_in_s * sr
        mask = torch.ones([1, n])
        if self.mask_type == "box":
            if self.start_inpainting_s is not None and self.end_inpainting_s is not None:
                mask[:, int(self.start_inpainting_s * sr): int(self.end_inpainting_s * sr)] = 0.
        elif self.mask_type == "random":
            width = int(self.mask_duration_s * sr)
            for _ in range(max(1, int(self.mask_percentage * n) // width)):
                start = torch.randint(0, mask.shape[1] - width, (1,))
                mask[:, start:start + width] = 0.
        elif self.mask_type == "periodic":
            width = int(self.mask_duration_s * sr)
            for start in range(0, mask.shape[1], int(self.interval_s * sr)):
                mask[:, start:min(start + width, mask.shape[1])] = 0.
        return mask

    def _mask_on(self, device):
        cache = self.__dict__.setdefault("_mask_cache", {})
        key = (str(device), self.mask.data_ptr(), self.mask._version)
        if cache.get("key") != key:
            cache["key"] = key
            cache["val"] = self.mask.to(device=device, dtype=torch.float32).reshape(-1).contiguous()
        return cache["val"]

    def transform(self, audio):
        return self._mel_db(audio)

    def forward(self, data, **kwargs):
        dev = _device_of(data)
        x = _as_f32_rows(data.detach().to(dev).reshape(-1, data.shape[-1]))
        mask = self._mask_on(dev)
        if mask.numel() != x.shape[1]:
            raise ValueError(f"mask has {mask.numel()} samples, data has {x.shape[1]}")
        y = torch.empty((x.shape[0], x.shape[1]), device=dev, dtype=torch.float32)
        _lib.call("dm_mask_apply", x.data_ptr(), x.stride(0), x.shape[1], x.shape[0], mask.data_ptr(), y.data_ptr(),
                  _lib.stream())
        return self._finish_forward(y.reshape(data.shape), data)

    def _fused(self, wav, measurement, space, want_grad, dwav=None):
        B, L = wav.shape
        mask = self._mask_on(wav.device)
        if mask.numel() != L:
            raise ValueError(f"mask has {mask.numel()} samples, waveform has {L}")
        if self._sigma() != 0.0:  # noisy A(x): materialise y = x*mask + sigma*n, then the generic stages
            y = self.forward(wav)
            ybar, pad, partial, nt = self._space_stage(y, measurement, space, want_grad)
        else:
            ybar, pad, partial, nt = self._space_stage(wav, measurement, space, want_grad, mask=mask)
        return self._fold_adjoint(ybar, pad, L, B, partial, nt, mask, want_grad, dwav, wav.dtype)


class PhaseRetrievalOperator(BaseOperator):
    """operator.py:136-171: A(x) = |STFT(x)| with a rectangular window; transform = clamp(mel of magnitude)."""

    window_kind = "rect"
    wave16 = True

    def __init__(self, n_fft=1024, hop_length=160, win_length=1024, noiser=None):
        if (n_fft, hop_length, win_length) != (1024, 160, 1024):
            raise NotImplementedError("the sm_100a STFT kernel is specialised for n_fft = win_length = 1024, hop 160 "
                                      "(configs/data/*.yaml); other sizes are not built")
        self.n_fft, self.hop_length, self.win_length = n_fft, hop_length, win_length
        self.noiser = noiser

    def transform(self, magnitude):
        """clamp(MelScale(64, 16000, 513)(magnitude.float()), +-80) -- operator.py:153-154.  The projection of an
        already-materialised magnitude is a (B*T, 513) x (513, 64) product with a 3 %-dense matrix; dm_mel_project
        evaluates it with the same banded table the fused kernel uses."""
        dev = _device_of(magnitude)
        lead = magnitude.shape[:-2]
        m = magnitude.detach().to(dev).float().reshape(-1, 513, magnitude.shape[-1]).contiguous()
        out = torch.empty((m.shape[0], 64, m.shape[2]), device=dev, dtype=torch.float32)
        _lib.call("dm_mel_project", self._tables(dev).ref, m.data_ptr(), m.shape[0], m.shape[2], 1, out.data_ptr(),
                  _lib.stream())
        return out.reshape(*lead, 64, m.shape[2]).to(magnitude.device)

    def forward(self, data, **kwargs):
        dev = _device_of(data)
        lead = data.shape[:-1]
        y = _as_f32_rows(data.detach().to(dev).reshape(-1, data.shape[-1]))
        mag = self._stft(_MODE_PHASE_WAV, y, out_rows=513, clamp=False)
        return self._finish_forward(mag.reshape(*lead, 513, mag.shape[-1]), data)

    def _fused(self, wav, measurement, space, want_grad, dwav=None):
        B, L = wav.shape
        sigma = self._sigma()
        T = 1 + L // _HOP
        noise = torch.randn((B, 513, T), device=wav.device, dtype=torch.float32) if sigma != 0.0 else None
        if space == "mel_spectrogram":
            ref = self._ref_mel(measurement).to(wav.device)
            ypbar, partial, nt = self._stft(_MODE_PHASE_MEL, wav, ref=ref, want_grad=want_grad, clamp=True,
                                            noise=noise, sigma=sigma)
        else:
            ref = measurement.to(wav.device).float().contiguous()
            if ref.shape[-2:] != (513, T):
                raise ValueError(f"measurement {tuple(ref.shape)} does not match |STFT| of the prediction (513, {T})")
            ypbar, partial, nt = self._stft(_MODE_PHASE_WAV, wav, ref=ref.reshape(-1, 513, T), want_grad=want_grad,
                                            clamp=False, noise=noise, sigma=sigma)
        return self._fold_adjoint(ypbar, 512, L, B, partial, nt, None, want_grad, dwav, wav.dtype)


class SuperResolutionOperator(BaseOperator):
    """operator.py:174-205: A(x) = torchaudio sinc resampling sample_rate -> sample_rate // scale."""

    def __init__(self, sample_rate, scale=10, noiser=None):
        self.resample_from, self.resample_to = sample_rate, sample_rate // scale
        self.kernel, self.width, self.orig, self.new = tables.sinc_resample_kernel(sample_rate, sample_rate // scale)
        self.sample_rate = 16000  # operator.py:181-189: wav2mel is built for 16 kHz whatever the input rate
        self.noiser = noiser

    def transform(self, audio):
        return self._mel_db(audio)

    wave16 = True  # scale 2 with 16-byte aligned rows: register-window kernels; otherwise the staged kernels convert

    def _kernel_on(self, device):
        cache = self.__dict__.setdefault("_k_cache", {})
        if str(device) not in cache:
            cache[str(device)] = self.kernel.to(device).contiguous()
        return cache[str(device)]

    def _resample(self, x, fill=None):
        """A(x); `fill`: a tensor the same launch zeroes on the way (the cotangent buffer of the chain)."""
        B, L = x.shape
        if self.kernel is None:  # orig == new: torchaudio returns the input unchanged
            if fill is not None:
                fill.zero_()
            return x.contiguous().clone()
        Ly = int(math.ceil(self.new * L / self.orig))
        k = self._kernel_on(x.device)
        y = torch.empty((B, Ly), device=x.device, dtype=torch.float32)
        if fill is None:
            _lib.call("dm_resample_fwd_io", x.data_ptr(), _lib.IO_DTYPES[x.dtype], x.stride(0), L, B, k.data_ptr(),
                      k.shape[0], k.shape[1], self.orig, self.width, y.data_ptr(), Ly, _lib.stream())
        else:
            _lib.call("dm_resample_fwd_fill_io", x.data_ptr(), _lib.IO_DTYPES[x.dtype], x.stride(0), L, B,
                      k.data_ptr(), k.shape[0], k.shape[1], self.orig, self.width, y.data_ptr(), Ly, fill.data_ptr(),
                      fill.numel(), _lib.stream())
        return y

    def forward(self, data, **kwargs):
        dev = _device_of(data)
        lead = data.shape[:-1]
        x = _as_f32_rows(data.detach().to(dev).reshape(-1, data.shape[-1]))
        y = self._resample(x)
        return self._finish_forward(y.reshape(*lead, y.shape[-1]), data)

    def _fir2_fusable(self, wav, space):
        """the shipped chain (scale 2, mel space, no measurement noise inside the step, fp32 rows the kernel can read
        with 128-bit loads) with the resampled signal computed inside the STFT kernel and never written to HBM.
        Opt-in (DM_STFT_FUSE_FIR=1): bit-identical, but measured SLOWER on B200 than the separate resampling launch
        (45.5 us against 33.5 + 9.2 us for 16 x 10 s clips: the window loads of the in-kernel FIR compete with the FFT
        exchanges for the same L1 / shared-memory pipe)."""
        return (space == "mel_spectrogram" and self.kernel is not None and (self.orig, self.new) == (2, 1)
                and tuple(self.kernel.shape) == (1, 28) and self.width == 13 and self._sigma() == 0.0
                and wav.dtype == torch.float32 and wav.shape[1] > 1024 and wav.data_ptr() % 16 == 0
                and wav.stride(0) % 4 == 0 and os.environ.get("DM_STFT_FUSE_FIR", "0") == "1")

    def _fused(self, wav, measurement, space, want_grad, dwav=None):
        B, L = wav.shape
        if self._fir2_fusable(wav, space):
            dev = wav.device
            Ly = (L + 1) // 2
            T = 1 + Ly // _HOP
            nf = _frames_per_tile(B, T, dev, max_nf=14)
            nt = math.ceil(T / nf)
            ref = self._ref_mel(measurement).to(dev)
            if tuple(ref.shape[1:]) != (64, T) or ref.shape[0] not in (1, B):
                raise ValueError(f"measurement transform has shape {tuple(ref.shape)}, prediction needs (1|{B}, 64, {T})")
            partial = torch.empty((B, nt), device=dev, dtype=torch.float32)
            ybar = torch.zeros((B, Ly + 1024), device=dev, dtype=torch.float32) if want_grad else None
            taps = self.__dict__.get("_taps_host")
            if taps is None:
                taps = self.__dict__["_taps_host"] = (C.c_float * 28)(*self.kernel[0].tolist())
            _lib.call("dm_stft_guidance_fir2", self._tables(dev).ref, int(self.clamp_transform), _HOP, wav.data_ptr(),
                      wav.stride(0), L, C.addressof(taps), B, ref.data_ptr(),
                      0 if ref.shape[0] == 1 else ref.stride(0), _lib.ptr(ybar), partial.data_ptr(), nf, _lib.stream())
            pad = 512
        else:
            ypbar = None
            if space == "mel_spectrogram" and want_grad:  # zeroed by the resampling launch: no fill kernel in the chain
                Ly = L if self.kernel is None else int(math.ceil(self.new * L / self.orig))
                ypbar = torch.empty((B, Ly + 1024), device=wav.device, dtype=torch.float32)
            y = self._resample(wav, fill=ypbar)
            if self._sigma() != 0.0:
                y = self._finish_forward(y, y)
            Ly = y.shape[1]
            ybar, pad, partial, nt = self._space_stage(y, measurement, space, want_grad, ypbar=ypbar)
        if not want_grad:
            return self._fold_adjoint(None, pad, Ly, B, partial, nt, None, False)
        if self.kernel is None:
            return self._fold_adjoint(ybar, pad, L, B, partial, nt, None, True, dwav)
        k = self._kernel_on(wav.device)
        loss = torch.empty((B,), device=wav.device, dtype=torch.float32)
        dwav = _grad_buffer(dwav, B, L, wav.device, wav.dtype)
        _lib.call("dm_resample_adjoint_io", ybar.data_ptr(), pad, Ly, B, partial.data_ptr(), nt, k.data_ptr(),
                  k.shape[0], k.shape[1], self.orig, self.width, dwav.data_ptr(), _lib.IO_DTYPES[wav.dtype],
                  dwav.stride(0), L, loss.data_ptr(), _lib.stream())
        return loss, dwav


class MusicDereverberationOperator(BaseOperator):
    """operator.py:208-250: A(x) = conv1d(x, ir, padding=K//2) with a random impulse response REDRAWN ON EVERY forward
    call from the global CPU generator (reference quirk kept: operator.py:246).  Evaluated by overlap-save FFT."""

    wave16 = True

    def __init__(self, ir_length=800, decay_factor=0.85, noiser=None):
        if ir_length > _RIR_MAX_TAPS:
            raise NotImplementedError(f"ir_length {ir_length} > {_RIR_MAX_TAPS} (one 8192-point FFT block)")
        self.ir_length = ir_length
        self.decay_factor = decay_factor
        self.noiser = noiser
        self.last_ir = None

    def transform(self, audio):
        return self._mel_db(audio)

    def generate_impulse_response(self, ir_length=800, decay_factor=0.85):
        """operator.py:238-242 -- host-side draw, same torch calls in the same order."""
        ir = torch.randn(ir_length)
        ir = torch.cumsum(ir, dim=0) * decay_factor
        ir /= ir.abs().max()
        return ir.unsqueeze(0)

    # ---- CUDA-graph support: the impulse response lives in a static device buffer the caller refreshes per step ----
    static_ir = None  # set by GraphedGuidedStep while it captures

    def _rir_tables(self, device):
        cache = self.__dict__.setdefault("_rir_cache", {})
        if str(device) not in cache:
            cache[str(device)] = (tables.twiddles(4096).to(device), tables.half_twiddles(8192).to(device))
        return cache[str(device)]

    def _spectrum(self, ir, device):
        tw, w = self._rir_tables(device)
        ir_d = ir.reshape(-1).to(device=device, dtype=torch.float32).contiguous()
        spec = torch.empty((4097, 2), device=device, dtype=torch.float32)
        _lib.call("dm_rir_spectrum", ir_d.data_ptr(), ir_d.numel(), tw.data_ptr(), w.data_ptr(), spec.data_ptr(),
                  _lib.stream())
        return spec, ir_d.numel()

    def _correlate(self, x, spec, K):
        B, L = x.shape
        tw, w = self._rir_tables(x.device)
        Ly = L + 2 * (K // 2) - K + 1
        y = torch.empty((B, Ly), device=x.device, dtype=torch.float32)
        _lib.call("dm_rir_correlate_io", x.data_ptr(), _lib.IO_DTYPES[x.dtype], x.stride(0), L, B, spec.data_ptr(), K,
                  tw.data_ptr(), w.data_ptr(), y.data_ptr(), Ly, _lib.stream())
        return y

    def forward(self, data, ir=None, **kwargs):
        dev = _device_of(data)
        if ir is None:
            ir = self.generate_impulse_response(ir_length=self.ir_length, decay_factor=self.decay_factor)
        self.last_ir = ir
        lead = data.shape[:-1]
        x = _as_f32_rows(data.detach().to(dev).reshape(-1, data.shape[-1]))
        spec, K = self._spectrum(ir, dev)
        y = self._correlate(x, spec, K)
        return self._finish_forward(y.reshape(*lead, y.shape[-1]), data)

    def _fused(self, wav, measurement, space, want_grad, dwav=None):
        B, L = wav.shape
        if self.static_ir is not None:
            ir = self.static_ir  # refreshed by GraphedGuidedStep before every replay (same host-side draw)
        else:
            ir = self.generate_impulse_response(ir_length=self.ir_length, decay_factor=self.decay_factor)
            self.last_ir = ir
        spec, K = self._spectrum(ir, wav.device)
        y = self._correlate(wav, spec, K)
        if self._sigma() != 0.0:
            y = self._finish_forward(y, y)
        Ly = y.shape[1]
        ybar, pad, partial, nt = self._space_stage(y, measurement, space, want_grad)
        if not want_grad:
            return self._fold_adjoint(None, pad, Ly, B, partial, nt, None, False)
        tw, w = self._rir_tables(wav.device)
        loss = torch.empty((B,), device=wav.device, dtype=torch.float32)
        dwav = _grad_buffer(dwav, B, L, wav.device, wav.dtype)
        _lib.call("dm_rir_adjoint_io", ybar.data_ptr(), pad, Ly, B, partial.data_ptr(), nt, spec.data_ptr(), K,
                  tw.data_ptr(), w.data_ptr(), dwav.data_ptr(), _lib.IO_DTYPES[wav.dtype], dwav.stride(0), L,
                  loss.data_ptr(), _lib.stream())
        return loss, dwav


class StyleGuidanceOperator(BaseOperator):
    """operator.py:253-271: forward = identity, transform = user-supplied CLAP gram matrix (a torch callable that
    stays in PyTorch; `get_gram_matrix` is not defined anywhere in the reference).  No fused kernel: the schedulers use
    the generic residual path for it."""

    def __init__(self, clap_model):
        self.clap_model = clap_model

    def transform(self, audio):
        return self.clap_model.get_gram_matrix(audio.float())

    def forward(self, data, **kwargs):
        return data

    fused_loss_and_grad = None  # no fused kernel: generic path

    def guidance_loss(self, wav, measurement, supervised_space="mel_spectrogram"):
        return generic_guidance_loss(self, wav, measurement, supervised_space)


def generic_guidance_loss(operator, wav, measurement, supervised_space):
    """Per-clip residual norm for operators without a fused kernel (user-defined torch callables): the reference's
    forward/transform/sub/norm chain (scheduling_dps.py:200-211) with the norm taken per clip."""
    pred = operator.forward(wav)
    if supervised_space == "wav_form":
        diff = measurement - pred
    elif supervised_space == "mel_spectrogram":
        diff = operator.transform(measurement) - operator.transform(pred)
    else:
        raise ValueError("supervised_space should be either 'wav_form' or 'mel_spectrogram")
    return torch.linalg.norm(diff.reshape(diff.shape[0], -1).float(), dim=1)
