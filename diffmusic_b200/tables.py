"""Host-side constant tables for the CUDA kernels.

The numeric constants are taken from the same torch / torchaudio calls the reference makes, so they are
bit-identical to what the reference multiplies with:

  * Hann window   : torchaudio MelSpectrogram -> Spectrogram.window = torch.hann_window(1024)  (operator.py:24-31)
  * mel filterbank: torchaudio.functional.melscale_fbanks(513, 0, sr/2, 64, sr, None, "htk")   (MelScale.fb;
                    operator.py:24-31 and 144-148)
  * sinc kernel   : torchaudio.functional._get_sinc_resample_kernel via T.Resample              (operator.py:180)

Index tables (band starts, per-bin band ids) are integer work done once at construction.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torchaudio

N_FFT = 1024
N_BINS = 513
N_MELS = 64
MEL_WSTRIDE = 42  # longest band of the 64-band HTK filterbank spans 41 bins


def hann_window():
    return torch.hann_window(N_FFT, periodic=True, dtype=torch.float32)


def rect_window():
    return torch.ones(N_FFT, dtype=torch.float32)


def mel_filterbank(sample_rate=16000):
    """(513, 64) fp32, exactly MelScale(n_mels=64, sample_rate, n_stft=513).fb"""
    return torchaudio.functional.melscale_fbanks(N_BINS, 0.0, float(sample_rate // 2), N_MELS, sample_rate, None,
                                                 "htk")


def inverse_mel_matrix(sample_rate=16000):
    """(64, 513) fp32 = W^T with W = fb (fb^T fb)^-1: torchaudio InverseMelScale(n_stft=513, n_mels=64, sample_rate)
    solves `lstsq(fb^T, mel, driver="gels")` per frame; fb has full column rank (cond(fb^T fb) = 32), so the solution of
    the under-determined system is the minimum-norm one, W mel.  Computed in float64 from the fp32 filterbank."""
    fb = mel_filterbank(sample_rate).double()
    w = fb @ torch.linalg.inv(fb.T @ fb)
    return w.T.contiguous().float()


def twiddles(n):
    """exp(-2 pi i m / n), m < n, computed in float64, stored as interleaved fp32 (n, 2)."""
    m = np.arange(n, dtype=np.float64)
    ang = -2.0 * math.pi * m / n
    return torch.from_numpy(np.stack([np.cos(ang), np.sin(ang)], axis=1).astype(np.float32))


def half_twiddles(n_real):
    """exp(-2 pi i k / n_real), k = 0..n_real/4, for the real-FFT unpack of an n_real-point transform."""
    k = np.arange(n_real // 4 + 1, dtype=np.float64)
    ang = -2.0 * math.pi * k / n_real
    return torch.from_numpy(np.stack([np.cos(ang), np.sin(ang)], axis=1).astype(np.float32))


def mel_tables(fb: torch.Tensor):
    """Sparse views of the triangular filterbank.

    Returns dict with
      mel_kstart (64,) int32, mel_klen (64,) int32, mel_w (MEL_WSTRIDE, 64) fp32   -- per band: consecutive bins,
                                                         stored TRANSPOSED (mel_w[i, m] = fb[kstart[m] + i, m]) so that
                                                         the 64 band-threads of a frame group read consecutive words
      bin_m0 (513,) int32, bin_w0 (513,), bin_w1 (513,) fp32                       -- per bin: <= 2 adjacent bands
    """
    fb = fb.detach().cpu().to(torch.float32)
    assert fb.shape == (N_BINS, N_MELS), fb.shape
    nz = fb != 0
    kstart = np.zeros(N_MELS, np.int32)
    klen = np.zeros(N_MELS, np.int32)
    w = np.zeros((N_MELS, MEL_WSTRIDE), np.float32)
    for m in range(N_MELS):
        ks = torch.nonzero(nz[:, m]).flatten().numpy()
        if ks.size == 0:
            continue
        k0, k1 = int(ks[0]), int(ks[-1])
        n = k1 - k0 + 1
        if n > MEL_WSTRIDE:
            raise ValueError(f"mel band {m} spans {n} bins > {MEL_WSTRIDE}")
        kstart[m], klen[m] = k0, n
        w[m, :n] = fb[k0:k1 + 1, m].numpy()  # zeros inside the span (none for triangles) are kept as zeros
    m0 = np.zeros(N_BINS, np.int32)
    w0 = np.zeros(N_BINS, np.float32)
    w1 = np.zeros(N_BINS, np.float32)
    for k in range(N_BINS):
        ms = torch.nonzero(nz[k]).flatten().numpy()
        if ms.size == 0:
            continue
        if ms.size > 2 or (ms.size == 2 and ms[1] != ms[0] + 1):
            raise ValueError(f"bin {k} touches non-adjacent mel bands {ms}")
        m0[k] = int(ms[0])
        w0[k] = float(fb[k, ms[0]])
        if ms.size == 2:
            w1[k] = float(fb[k, ms[1]])
    return dict(mel_kstart=torch.from_numpy(kstart), mel_klen=torch.from_numpy(klen),
                mel_w=torch.from_numpy(np.ascontiguousarray(w.T)),
                bin_m0=torch.from_numpy(m0), bin_w0=torch.from_numpy(w0), bin_w1=torch.from_numpy(w1))


def sinc_resample_kernel(orig_freq, new_freq):
    """(kernel (new, taps) fp32, width, orig, new) exactly as torchaudio.transforms.Resample builds them."""
    rs = torchaudio.transforms.Resample(orig_freq=orig_freq, new_freq=new_freq)
    g = math.gcd(int(orig_freq), int(new_freq))
    orig, new = int(orig_freq) // g, int(new_freq) // g
    if orig == new:
        return None, 0, orig, new
    return rs.kernel[:, 0, :].contiguous().to(torch.float32), int(rs.width), orig, new
