"""`mel_spectrogram_to_waveform_with_phase` of the reference pipelines (diffmusic/pipelines/pipeline_musicldm.py:263-301,
plpeline_audioldm2.py:681; the export path of run.py:343) on the GPU, SURVEY.md 8f rank 3:

    mel (B, 1, T, 64) -> InverseMelScale(513, 64, 16000) -> * exp(1j * original_phase) -> torch.istft(1024, hop, 1024)
    -> clip / zero-pad to original_waveform_length

One fused chain (csrc/istft.cu): the minimum-norm inverse mel projection is a fixed 513 x 64 matrix applied per tile,
the complex spectrogram lives in shared memory only, the inverse FFT runs two frames per 64-thread group on the
frame-pair pipeline of the STFT guidance kernel, overlap-add + envelope division + trimming finish the waveform.
There is no CPU implementation here.
"""
from __future__ import annotations

import torch

from . import _lib, tables
from .operators import _DeviceTables

_TABLES = {}


def _tables(dev):
    key = str(dev)
    if key not in _TABLES:
        _TABLES[key] = (_DeviceTables(dev, tables.rect_window(), 16000), tables.inverse_mel_matrix(16000).to(dev))
    return _TABLES[key]


def mel_spectrogram_to_waveform_with_phase(mel_spectrogram, original_phase, n_fft=1024, hop_length=160, win_length=1024,
                                           original_waveform_length=0):
    """Same arguments and result as the pipeline method: `mel_spectrogram` (B, 1, T, n_mels) or (B, T, n_mels),
    `original_phase` (1, 513, T) (shared by the batch) or (B, 513, T); returns (B, hop (T - 1)) fp32, clipped or
    zero-padded to `original_waveform_length` when that is positive."""
    if (n_fft, win_length) != (1024, 1024):
        raise NotImplementedError("the STFT kernels are built for n_fft = win_length = 1024")
    if hop_length <= 0 or hop_length > 1024 or hop_length % 2:
        raise NotImplementedError("hop_length must be even and <= n_fft")
    _lib.require_cuda(mel_spectrogram)
    dev = mel_spectrogram.device
    mel = mel_spectrogram.squeeze(1) if mel_spectrogram.dim() == 4 else mel_spectrogram  # (B, T, n_mels)
    if mel.dim() != 3 or mel.shape[-1] != tables.N_MELS:
        raise ValueError(f"Expected an input with {tables.N_MELS} mel bins. Found: {mel.shape[-1]}")
    phase = original_phase.squeeze(0)
    if phase.dtype.is_complex:
        raise ValueError("original_phase must be real (angles in radians)")
    mel = mel.to(device=dev, dtype=torch.float32)
    phase = phase.to(device=dev, dtype=torch.float32).contiguous()
    B, T = mel.shape[0], mel.shape[1]
    if phase.dim() == 2:
        phase_bstride = 0
    elif phase.dim() == 3 and phase.shape[0] == B:
        phase_bstride = phase.stride(0)
    else:
        raise ValueError(f"original_phase {tuple(original_phase.shape)} does not match a batch of {B}")
    if tuple(phase.shape[-2:]) != (tables.N_BINS, T):
        raise RuntimeError(f"The size of tensor a ({T}) must match the size of tensor b ({phase.shape[-1]}) at "
                           f"non-singleton dimension 2")
    if T < 2:
        raise RuntimeError("istft needs at least two frames")
    tab, winv_t = _tables(dev)
    n = hop_length * (T - 1)
    out_len = int(original_waveform_length) if original_waveform_length > 0 else n
    lib = _lib.load()
    ola = torch.empty(int(lib.dm_istft_workspace_floats(B, T, hop_length)), device=dev, dtype=torch.float32)
    out = torch.empty((B, out_len), device=dev, dtype=torch.float32)
    _lib.call("dm_istft_mel_phase", tab.ref, winv_t.data_ptr(), mel.data_ptr(), mel.stride(0), mel.stride(2),
              mel.stride(1), phase.data_ptr(), phase_bstride, B, T, hop_length, ola.data_ptr(), out.data_ptr(), out_len,
              _lib.stream())
    return out


def waveform_to_spectrogram(waveform, n_fft=1024, hop_length=160, win_length=1024):
    """diffmusic/utils.py:11-20 (run.py:305 takes the degraded clip's phase from it): `torch.stft(waveform, n_fft,
    hop_length, win_length, return_complex=True)` -- no window given, i.e. rectangular; centred, reflect padding -- and
    its `(abs, angle)`, each (B, 513, 1 + L // hop) fp32.  One launch, the complex spectrogram is never materialised."""
    if (n_fft, win_length) != (1024, 1024):
        raise NotImplementedError("the STFT kernels are built for n_fft = win_length = 1024")
    if hop_length <= 0 or hop_length > 1024 or hop_length % 2:
        raise NotImplementedError("hop_length must be even and <= n_fft")
    _lib.require_cuda(waveform)
    squeeze = waveform.dim() == 1
    wav = (waveform[None] if squeeze else waveform).to(torch.float32)
    if wav.dim() != 2:
        raise RuntimeError(f"stft: expected a 1D or 2D tensor, got {waveform.dim()}D")
    if wav.stride(1) != 1:
        wav = wav.contiguous()
    B, L = wav.shape
    if L <= 512:
        raise RuntimeError(f"Argument #4: Padding size should be less than the corresponding input dimension, but got: "
                           f"padding (512, 512) at dimension 2 of input {[1, B, L]}")
    dev = wav.device
    T = 1 + L // hop_length
    tab, _ = _tables(dev)
    mag = torch.empty((B, tables.N_BINS, T), device=dev, dtype=torch.float32)
    phase = torch.empty_like(mag)
    _lib.call("dm_stft_spectrogram", tab.ref, wav.data_ptr(), wav.stride(0), L, B, hop_length, mag.data_ptr(),
              phase.data_ptr(), _lib.stream())
    return (mag[0], phase[0]) if squeeze else (mag, phase)
