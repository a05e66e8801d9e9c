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


def warp_image(window: torch.Tensor, fb: torch.Tensor):
    """Shared-memory image of the constant tables of the warp-per-frame-pair STFT kernel (csrc/stft_warp.cu), built once
    on the host so that a CTA fetches it with ONE bulk copy instead of recomputing it.  Returns (image fp32 (n,), na, nb).

    Sections (float offsets, each a multiple of 4 floats; the kernel derives them from na / nb):
      win2  [512][2]        (window[n], window[n + 512]) / 2   -- the 1/2 of the two-real-frames split rides on the window
      tw4   [16][32][4]     (w^(2m), w^(2m+1)) for w = exp(-2 pi i lane / 1024): the twiddles between the two DFT-32 passes
      melp  [na + nb][32][2] filterbank weights of lane l for bin PAIRS: rows < na belong to its short band ma[l] (pairs
                            pa0[l] + i), rows >= na to its long band mb[l] (pairs pb0[l] + i - na); zero outside the band,
                            so every lane runs the same na + nb iterations with 128-bit loads of (P_A, P_B) of two bins
      lanek [4][32] int32   pa0, pb0 (first bin pair of lane l's two row blocks), ma, mb (its two mel bands)
      binw  [514][2]        (fb[k, m0], fb[k, m0 + 1]) per bin;   binm [528] uint8  m0 per bin
    """
    mt = mel_tables(fb)
    fbn = fb.detach().cpu().to(torch.float32).numpy()
    k0, kl = mt["mel_kstart"].numpy(), mt["mel_klen"].numpy()
    win = window.detach().cpu().to(torch.float32).numpy()
    win2 = np.stack([0.5 * win[:512], 0.5 * win[512:]], axis=1).astype(np.float32)
    lane = np.arange(32, dtype=np.float64)
    tw4 = np.zeros((16, 32, 4), np.float32)
    for m in range(16):
        for j, e in enumerate((2 * m, 2 * m + 1)):
            ang = -2.0 * math.pi * lane * e / 1024.0
            tw4[m, :, 2 * j] = np.cos(ang)
            tw4[m, :, 2 * j + 1] = np.sin(ang)

    def pairs(m):
        a, b = int(k0[m]), int(k0[m] + kl[m] - 1)
        return a >> 1, (b >> 1) - (a >> 1) + 1

    # Band -> lane assignment.  Lane l sums one SHORT band (0..31) in `na` rows and one LONG band (32..63) in `nb` rows;
    # every lane runs all rows (zero weights outside its band), so which bands a lane takes is free.  The 128-bit loads
    # of row i hit bin pair p0[l] + i: a quarter-warp (8 consecutive lanes) is conflict-free iff its eight p0 are
    # distinct mod 8.  A band shorter than the row count may start its window up to (rows - count) pairs early, which
    # gives every band a set of reachable residues; bands are matched to residues (capacity 4 = one per quarter-warp)
    # by augmenting paths, and the band with residue r of quarter-warp o sits in lane 8 o + r.
    def assign(bands, rows):
        reach = {}
        for m in bands:
            p0, cnt = pairs(m)
            lo = max(0, p0 - (rows - cnt))
            hi = min(p0, 257 - rows)
            assert lo <= hi, (m, p0, cnt, rows)
            reach[m] = {}
            for start in range(hi, lo - 1, -1):      # prefer the latest start (fewest padded rows in front)
                reach[m].setdefault(start % 8, start)
        owner = {r: [] for r in range(8)}

        def place(m, seen):
            for r in reach[m]:
                if r in seen:
                    continue
                seen.add(r)
                if len(owner[r]) < 4:
                    owner[r].append(m)
                    return True
                for j, other in enumerate(owner[r]):
                    if place(other, seen):
                        owner[r][j] = m
                        return True
            return False

        for m in sorted(bands, key=lambda m: len(reach[m])):
            if not place(m, set()):
                return None
        lanes = [None] * 32
        for r in range(8):
            for o, m in enumerate(owner[r]):
                lanes[8 * o + r] = (m, reach[m][r])
        return lanes

    na = max(pairs(m)[1] for m in range(32))
    nb = max(pairs(m)[1] for m in range(32, 64))
    la, lb = assign(range(32), na), assign(range(32, 64), nb)
    if la is None or lb is None:  # no conflict-free assignment (other filterbanks): natural order, still correct
        la = [(l, min(pairs(l)[0], 257 - na)) for l in range(32)]
        lb = [(63 - l, min(pairs(63 - l)[0], 257 - nb)) for l in range(32)]
    melp = np.zeros((na + nb, 32, 2), np.float32)
    chk = np.zeros_like(fbn)
    for l in range(32):
        for (m, p0), base, n in ((la[l], 0, na), (lb[l], na, nb)):
            assert 0 <= p0 and 2 * (p0 + n - 1) + 1 <= 513, (l, m, p0, n)  # padded reads stay inside P[0 .. 513]
            for i in range(n):
                for h in range(2):
                    k = 2 * (p0 + i) + h
                    if k < N_BINS:
                        melp[base + i, l, h] = fbn[k, m]
                        chk[k, m] = fbn[k, m]
            assert int(k0[m]) >= 2 * p0 and int(k0[m] + kl[m] - 1) <= 2 * (p0 + n - 1) + 1, (l, m)
    assert np.array_equal(chk, fbn)  # the pair tables reproduce the filterbank exactly
    assert sorted(m for m, _ in la) == list(range(32)) and sorted(m for m, _ in lb) == list(range(32, 64))
    lanek = np.array([[p for _, p in la], [p for _, p in lb], [m for m, _ in la], [m for m, _ in lb]], np.int32)
    binw = np.zeros((514, 2), np.float32)
    binw[:N_BINS, 0] = mt["bin_w0"].numpy()
    binw[:N_BINS, 1] = mt["bin_w1"].numpy()
    binm = np.zeros(528, np.uint8)
    binm[:N_BINS] = mt["bin_m0"].numpy().astype(np.uint8)
    parts = [win2.ravel(), tw4.ravel(), melp.ravel(), lanek.ravel().view(np.float32), binw.ravel(),
             binm.view(np.float32)]
    for q in parts:
        assert q.size % 4 == 0
    return torch.from_numpy(np.concatenate(parts).copy()), int(na), int(nb)


def sinc_resample_kernel(orig_freq, new_freq):
    """(kernel (new, taps) fp32, width, orig, new) exactly as torchaudio.transforms.Resample builds them."""
    rs = torchaudio.transforms.Resample(orig_freq=orig_freq, new_freq=new_freq)
    g = math.gcd(int(orig_freq), int(new_freq))
    orig, new = int(orig_freq) // g, int(new_freq) // g
    if orig == new:
        return None, 0, orig, new
    return rs.kernel[:, 0, :].contiguous().to(torch.float32), int(rs.width), orig, new
