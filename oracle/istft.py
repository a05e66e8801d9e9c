"""Restatement of `mel_spectrogram_to_waveform_with_phase` (diffmusic/pipelines/pipeline_musicldm.py:263-301, identical
in plpeline_audioldm2.py:681) -- TEST INFRASTRUCTURE (see oracle/__init__.py).

Third-party arithmetic on the path: torchaudio.transforms.InverseMelScale (driver "gels") and torch.istft; both are in
this image, so the restatement is pinned twice: against the REFERENCE function itself, extracted from the reference source
and run in the build container (tests/golden/istft.npz, made by tests/golden/make_istft_golden.py), and against the
live libraries (tests/test_oracle_vs_golden.py).

Published algorithms restated in float64:
  * InverseMelScale.forward: relu(lstsq(fb^T, mel).solution) per frame; fb (513 x 64, fp32 triangular HTK filterbank,
    f_max = sr // 2) has full column rank, so `gels` on the under-determined system returns the minimum-norm solution
    fb (fb^T fb)^-1 mel.
  * torch.istft(n_fft, hop, win_length = n_fft, window=None, center=True, normalized=False, onesided, length=None):
    frames = irfft(spec, n = n_fft) (imaginary parts of the DC / Nyquist bins ignored, 1/n_fft scaling), rectangular
    window, overlap-add to n_fft + hop (T - 1) samples, divided by the overlap-added squared window (= number of frames
    covering a sample), samples [n_fft/2, end - n_fft/2) kept.
"""
from __future__ import annotations

import numpy as np
import torchaudio


def filterbank(n_stft=513, n_mels=64, sample_rate=16000):
    return torchaudio.functional.melscale_fbanks(n_stft, 0.0, float(sample_rate // 2), n_mels, sample_rate, None,
                                                 "htk").numpy().astype(np.float64)


def inverse_mel_scale(mel, n_stft=513, sample_rate=16000):
    """mel (..., n_mels, T) -> (..., n_stft, T): pipeline_musicldm.py:277-281."""
    mel = np.asarray(mel, np.float64)
    fb = filterbank(n_stft, mel.shape[-2], sample_rate)
    w = fb @ np.linalg.inv(fb.T @ fb)
    return np.maximum(np.einsum("km,...mt->...kt", w, mel), 0.0)


def istft_rect(spec, n_fft=1024, hop=160):
    """spec (B, n_fft/2 + 1, T) complex -> (B, hop (T - 1)): pipeline_musicldm.py:283-288."""
    spec = np.asarray(spec, np.complex128)
    B, _, T = spec.shape
    frames = np.fft.irfft(spec, n=n_fft, axis=1)  # (B, n_fft, T)
    n = n_fft + hop * (T - 1)
    y = np.zeros((B, n))
    env = np.zeros(n)
    for t in range(T):
        y[:, t * hop:t * hop + n_fft] += frames[:, :, t]
        env[t * hop:t * hop + n_fft] += 1.0
    half = n_fft // 2
    return y[:, half:n - half] / env[half:n - half]


def mel_spectrogram_to_waveform_with_phase(mel_spectrogram, original_phase, n_fft=1024, hop_length=160,
                                           win_length=1024, original_waveform_length=0):
    """pipeline_musicldm.py:263-301 line by line (win_length = n_fft)."""
    assert win_length == n_fft
    mel = np.asarray(mel_spectrogram, np.float64)
    mel = np.swapaxes(mel[:, 0] if mel.ndim == 4 else mel, 1, 2)  # squeeze(1).permute(0, 2, 1): (B, n_mels, T)
    phase = np.asarray(original_phase, np.float64)
    if phase.shape[0] == 1:
        phase = phase[0]
    lin = inverse_mel_scale(mel, n_fft // 2 + 1)
    wav = istft_rect(lin * np.exp(1j * phase), n_fft, hop_length)
    if original_waveform_length > 0:
        if wav.shape[-1] > original_waveform_length:
            wav = wav[..., :original_waveform_length]
        elif wav.shape[-1] < original_waveform_length:
            wav = np.pad(wav, ((0, 0), (0, original_waveform_length - wav.shape[-1])))
    return wav


def waveform_to_spectrogram(waveform, n_fft=1024, hop_length=160, win_length=1024):
    """diffmusic/utils.py:11-20 restated in float64: rectangular window (none is passed), centred frames over the
    reflect-padded signal, one-sided DFT; returns (abs, angle), each (B, n_fft/2 + 1, 1 + L // hop)."""
    assert win_length == n_fft
    x = np.asarray(waveform, np.float64)
    xp = np.pad(x, ((0, 0), (n_fft // 2, n_fft // 2)), mode="reflect")
    T = 1 + x.shape[1] // hop_length
    frames = np.stack([xp[:, t * hop_length:t * hop_length + n_fft] for t in range(T)], axis=2)  # (B, n_fft, T)
    spec = np.fft.rfft(frames, axis=1)
    return np.abs(spec), np.angle(spec)
