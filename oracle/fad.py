"""NumPy restatement of the FAD embedding statistics -- TEST INFRASTRUCTURE (see oracle/__init__.py).

fadtk/fad.py:41-47 (`calc_embd_statistics`) and fadtk/utils.py:13-46 (`_process_file`,
`calculate_embd_statistics_online`).  fadtk/utils.py cannot be imported here (hypy_utils is absent); its arithmetic
is np.mean / np.cov (float64) and a pairwise (Chan) merge, restated below on in-memory arrays instead of .npy files.
"""
from __future__ import annotations

import numpy as np


def calc_embd_statistics(embd):
    """fadtk/fad.py:41-47: (np.mean(axis 0) [dtype of the input -- fp16 for cached embeddings], np.cov fp64)."""
    assert embd.shape[0] >= 2, "FAD requires at least two embedding window frames"
    return np.mean(embd, axis=0), np.cov(embd, rowvar=False)


def process_one(embd):
    """fadtk/utils.py:13-16 on an array instead of a file: (mean, cov * (n - 1), n)."""
    n = embd.shape[0]
    return np.mean(embd, axis=0), np.cov(embd, rowvar=False) * (n - 1), n


def embd_statistics_online(arrays):
    """fadtk/utils.py:19-46: Chan merge of per-file (mean, scatter, n)."""
    assert len(arrays) > 0, "No files provided"
    d = arrays[0].shape[-1]
    mu = np.zeros(d)
    S = np.zeros((d, d))
    n = 0
    for a in arrays:
        _mu, _S, _n = process_one(a)
        delta = _mu - mu
        mu += _n / (n + _n) * delta
        S += _S + delta[:, None] * delta[None, :] * n * _n / (n + _n)
        n += _n
    if n < 2:
        return mu, np.zeros_like(S)
    return mu, S / (n - 1)


def moments_to_stats(n, sx, sxx):
    """What the product's all-reduced raw moments (n, sum x, sum x x^T; float64) must finalise to."""
    mu = sx / n
    if n < 2:
        return mu, np.zeros_like(sxx)
    return mu, (sxx - n * np.outer(mu, mu)) / (n - 1)
