"""NumPy restatement of the FAD embedding statistics -- TEST INFRASTRUCTURE (see oracle/__init__.py).

fadtk/fad.py:41-47 (`calc_embd_statistics`) and fadtk/utils.py:13-46 (`_process_file`,
`calculate_embd_statistics_online`).  fadtk/utils.py cannot be imported here (hypy_utils is absent); its arithmetic
is np.mean / np.cov (float64) and a pairwise (Chan) merge, restated below on in-memory arrays instead of .npy files.

`calc_frechet_distance` / `score_inf` restate fadtk/fad.py:50-119, 303-350 with the same SciPy / NumPy calls.  PINNED:
tests/golden/make_fad_golden.py cuts the reference's own functions (calc_embd_statistics, calc_frechet_distance,
FrechetAudioDistanceTK.score_inf, _process_file, calculate_embd_statistics_online) out of fadtk/fad.py and fadtk/utils.py
with `ast` and runs them unmodified on seeded fp16 embeddings; tests/test_oracle_vs_golden.py holds every function of this
file to those outputs (tests/golden/fad.npz), next to the closed forms (identical Gaussians -> 0, commuting covariances).
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


def calc_frechet_distance(mu1, cov1, mu2, cov2, eps=1e-6):
    """fadtk/fad.py:50-119, eigenvalue method (the value the reference returns): d^2 = |mu1-mu2|^2 + tr C1 + tr C2
    - 2 tr((V sqrt(D)) V^-1) with D, V = eig(C1 C2).  The scipy.linalg.sqrtm call of the reference only feeds a log
    message and is omitted."""
    from scipy import linalg
    mu1, mu2 = np.atleast_1d(mu1), np.atleast_1d(mu2)
    cov1, cov2 = np.atleast_2d(cov1), np.atleast_2d(cov2)
    assert mu1.shape == mu2.shape and cov1.shape == cov2.shape
    diff = mu1 - mu2
    D, V = linalg.eig(cov1.dot(cov2))
    covmean = (V * np.emath.sqrt(D)) @ linalg.inv(V)
    if not np.isfinite(covmean).all():
        offset = np.eye(cov1.shape[0]) * eps
        covmean = linalg.sqrtm((cov1 + offset).dot(cov2 + offset))
    if np.iscomplexobj(covmean):
        if not np.allclose(np.diagonal(covmean).imag, 0, atol=1e-3):
            raise ValueError('Imaginary component {}'.format(np.max(np.abs(covmean.imag))))
        covmean = covmean.real
    return diff.dot(diff) + np.trace(cov1) + np.trace(cov2) - 2 * np.trace(covmean)


def score_inf(mu_base, cov_base, embeds, steps=25, min_n=500):
    """fadtk/fad.py:303-350 on an in-memory (N, d) array: (score, slope, r2, points)."""
    max_n = len(embeds)
    ns = [int(n) for n in np.linspace(min_n, max_n, steps)]
    results = []
    for n in ns:
        indices = np.random.choice(embeds.shape[0], size=n, replace=True)
        mu_eval, cov_eval = calc_embd_statistics(embeds[indices])
        results.append([n, calc_frechet_distance(mu_base, cov_base, mu_eval, cov_eval)])
    ys = np.array(results)
    xs = 1 / np.array(ns)
    slope, intercept = np.polyfit(xs, ys[:, 1], 1)
    r2 = 1 - np.sum((ys[:, 1] - (slope * xs + intercept)) ** 2) / np.sum((ys[:, 1] - np.mean(ys[:, 1])) ** 2)
    return intercept, slope, r2, results
