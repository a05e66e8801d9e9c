"""Evaluation metrics of diffmusic/metrics on the GPU: LogSpectralDistance (lsd.py:17-40) and
MeanSquaredError (mse.py:9-29), same class names, constructor arguments and `score` signatures.

LSD runs on the frame-pair STFT pipeline (csrc/metrics.cu): background and eval frames go through ONE FFT as a pair and
the per-frame distance is reduced on chip; nothing but the (B, T) per-frame distances leaves the SM.
The reference calls `librosa.stft(y, n_fft, hop_length)`: Hann (periodic) window of n_fft samples, centred frames.
librosa is unpinned in requirements.txt and absent from this image, so its padding default cannot be probed here: librosa
>= 0.10 pads with zeros (`pad_mode="constant"`, the default used below), older releases reflect (`pad_mode="reflect"`).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib, tables
from .operators import _DeviceTables


def _device():
    if not torch.cuda.is_available():
        raise _lib.DiffMusicB200Error("diffmusic_b200.metrics needs CUDA (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def _rows(a, dev):
    t = torch.as_tensor(np.asarray(a) if not isinstance(a, torch.Tensor) else a)
    if t.dim() == 1:
        t = t[None]
    return t.to(device=dev, dtype=torch.float32).contiguous()


class LogSpectralDistance:
    _tab_cache = {}

    def __init__(self, sample_rate=16000, n_fft=1024, hop_length=160, eps=1e-10, pad_mode="constant"):
        if n_fft != 1024:
            raise NotImplementedError("the STFT kernels are built for n_fft = 1024 (eval.py:124-129)")
        if hop_length <= 0 or hop_length > 1024 or hop_length % 2:
            raise NotImplementedError("hop_length must be even and <= n_fft")
        if pad_mode not in ("constant", "reflect"):
            raise ValueError("pad_mode must be 'constant' or 'reflect'")
        self.sample_rate, self.n_fft, self.hop_length, self.eps, self.pad_mode = sample_rate, n_fft, hop_length, eps, pad_mode

    def _tables(self, dev):
        key = str(dev)
        if key not in self._tab_cache:
            self._tab_cache[key] = _DeviceTables(dev, tables.hann_window(), 16000)
        return self._tab_cache[key]

    def frame_distances(self, audio_background, audio_eval):
        """(B, T) per-frame sqrt(mean_k (log10|X_bg| - log10|X_eval|)^2) as a device tensor."""
        dev = _device()
        bg, ev = _rows(audio_background, dev), _rows(audio_eval, dev)
        if bg.shape != ev.shape:
            raise ValueError(f"operands could not be broadcast together with shapes {tuple(bg.shape)} {tuple(ev.shape)}")
        B, L = bg.shape
        T = 1 + L // self.hop_length
        out = torch.empty((B, T), device=dev, dtype=torch.float32)
        _lib.call("dm_lsd_frames", self._tables(dev).ref, bg.data_ptr(), bg.stride(0), ev.data_ptr(), ev.stride(0), L, B,
                  self.hop_length, int(self.pad_mode == "reflect"), 0, float(self.eps), out.data_ptr(), _lib.stream())
        return out

    def score(self, audio_background, audio_eval, output_mean=True):
        lsd_score = self.frame_distances(audio_background, audio_eval).mean(dim=1)
        if output_mean:
            return float(lsd_score.mean().item())
        return lsd_score.cpu().numpy()


class MeanSquaredError:
    def __init__(self, reduction='mean'):
        assert reduction in ['mean', 'sum'], "reduction must be 'mean' or 'sum'"
        self.reduction = reduction

    @staticmethod
    def per_clip(ref, est):
        """(B,) mean squared difference per clip over the common length, device tensor."""
        dev = _device()
        r, e = _rows(ref, dev), _rows(est, dev)
        if r.shape[0] != e.shape[0]:
            raise ValueError("background and eval hold different numbers of clips")
        n = min(r.shape[1], e.shape[1])
        B = r.shape[0]
        lib = _lib.load()
        lib.dm_mse_num_chunks.restype = C.c_longlong
        partial = torch.empty((B, int(lib.dm_mse_num_chunks(C.c_longlong(n)))), device=dev, dtype=torch.float64)
        out = torch.empty(B, device=dev, dtype=torch.float32)
        _lib.call("dm_mse", r.data_ptr(), r.stride(0), e.data_ptr(), e.stride(0), n, B, partial.data_ptr(),
                  out.data_ptr(), _lib.stream())
        return out

    def score(self, audio_background, audio_eval):
        try:
            same = len({len(x) for x in audio_background} | {len(x) for x in audio_eval}) == 1
        except TypeError:
            same = True
        if same:
            mse_scores = self.per_clip(audio_background, audio_eval)
        else:  # ragged lists: the reference truncates every pair to its common length (mse.py:19-22)
            mse_scores = torch.cat([self.per_clip(r, e) for r, e in zip(audio_background, audio_eval)])
        v = mse_scores.mean() if self.reduction == 'mean' else mse_scores.sum()
        return float(v.item())
