"""Restatement of diffmusic/metrics/lsd.py:17-40 and mse.py:9-29 -- TEST INFRASTRUCTURE (see oracle/__init__.py).

lsd.py calls librosa.stft(y, n_fft=..., hop_length=...), a third-party dependency that is unpinned in requirements.txt
and absent from this image (PARITY UNPINNED for that call).  Its published algorithm: frames of n_fft samples, centred
(n_fft // 2 samples of padding on both sides, zeros for librosa >= 0.10 / reflection before), periodic Hann window,
complex64 output for float32 input, T = 1 + L // hop frames.  torch.stft with the same window / padding is that
computation; everything after it is the reference's NumPy code line by line.
"""
from __future__ import annotations

import numpy as np
import torch


def _stft_mag(y, n_fft, hop, pad_mode):
    y = torch.as_tensor(np.asarray(y, dtype=np.float32))
    spec = torch.stft(y, n_fft, hop_length=hop, win_length=n_fft, window=torch.hann_window(n_fft), center=True,
                      pad_mode=pad_mode, return_complex=True)
    return np.abs(spec.numpy())


def lsd_score(audio_background, audio_eval, n_fft=1024, hop_length=160, eps=1e-10, output_mean=True,
              pad_mode="constant"):
    """lsd.py:17-40."""
    audio_background = np.array(audio_background)
    audio_eval = np.array(audio_eval)
    audio_eval = np.nan_to_num(audio_eval, nan=0.0, posinf=1.0, neginf=-1.0)
    background_spectrogram = _stft_mag(audio_background, n_fft, hop_length, pad_mode)
    eval_spectrogram = _stft_mag(audio_eval, n_fft, hop_length, pad_mode)
    log_background = np.log10(background_spectrogram + eps)
    log_eval = np.log10(eval_spectrogram + eps)
    squared_diff = (log_background - log_eval) ** 2
    lsd_per_frame = np.sqrt(np.mean(squared_diff, axis=1))
    lsd = np.mean(lsd_per_frame, axis=1)
    return lsd.mean() if output_mean else lsd


def mse_score(audio_background, audio_eval, reduction="mean"):
    """mse.py:9-29."""
    audio_background = np.array(audio_background, dtype=np.float32)
    audio_eval = np.array(audio_eval, dtype=np.float32)
    audio_eval = np.nan_to_num(audio_eval, nan=0.0, posinf=1.0, neginf=-1.0)
    audio_background = np.nan_to_num(audio_background, nan=0.0, posinf=1.0, neginf=-1.0)
    scores = []
    for ref, est in zip(audio_background, audio_eval):
        n = min(len(ref), len(est))
        scores.append(np.mean((ref[:n] - est[:n]) ** 2))
    scores = np.array(scores)
    return scores.mean() if reduction == "mean" else scores.sum()
