"""Fréchet distance on the GPU: device time of dm_frechet_distance per embedding width, next to the
host computing what fadtk/fad.py:50-119 computes (eigvals of the d x d product, all host threads).

    python tools/frechet_bench.py [--dims 128 512 768] > gpurun_out/frechet_bench.json
    DM_JACOBI_GRAPH=0 python tools/frechet_bench.py ...     # plain launch loop instead of graph-replayed sweeps

Inputs: covariances of two synthetic embedding sets (4096 x d, different means / scales), float64.  Timing: CUDA events
around `iters` back-to-back solves after one warm-up, no host synchronisation inside."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffmusic_b200 import fad  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dims", type=int, nargs="+", default=[128, 512, 768])
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--cpu", type=int, default=1)
    ap.add_argument("--inf", type=int, default=0, help="also time score_inf at this embedding width")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    rows = []
    for d in a.dims:
        rng = np.random.default_rng(d)
        x1 = rng.standard_normal((4096, d)) * (0.5 + rng.random(d)) + 0.1
        x2 = rng.standard_normal((4096, d)) * (0.4 + rng.random(d)) - 0.05
        mu1, mu2, c1, c2 = x1.mean(0), x2.mean(0), np.cov(x1, rowvar=False), np.cov(x2, rowvar=False)
        args = [torch.as_tensor(v).to(dev) for v in (mu1, c1, mu2, c2)]
        out = fad.frechet_distance_device(*args)
        torch.cuda.synchronize()
        s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h0 = time.perf_counter()
        s.record()
        for _ in range(a.iters):
            out = fad.frechet_distance_device(*args)
        t.record()
        h1 = time.perf_counter()
        torch.cuda.synchronize()
        o = out.cpu().numpy()
        row = {"d": d, "gpu_ms": s.elapsed_time(t) / a.iters, "host_enqueue_ms": (h1 - h0) * 1e3 / a.iters,
               "value": float(o[0]), "sweeps": [int(o[2]), int(o[3])]}
        if a.cpu:
            c0 = time.perf_counter()
            ev = np.linalg.eigvals(c1 @ c2)
            want = float(((mu1 - mu2) ** 2).sum() + np.trace(c1) + np.trace(c2) - 2.0 * np.sqrt(np.abs(ev.real)).sum())
            row["cpu_eigvals_ms"] = (time.perf_counter() - c0) * 1e3
            row["cpu_threads"] = torch.get_num_threads()
            row["rel_diff_vs_cpu"] = abs(row["value"] - want) / abs(want)
        rows.append(row)
    inf = None
    if a.inf:  # FAD-inf (fadtk/fad.py:303-350): 25 bootstrap points = 25 x (gather + moments + Frechet distance)
        d, n = a.inf, 8000
        rng = np.random.default_rng(3)
        base = rng.standard_normal((4096, d)) * 0.6 + 0.2
        emb = (rng.standard_normal((n, d)) * 0.5 + 0.25).astype(np.float16)
        mu_b, cov_b = base.mean(0), np.cov(base, rowvar=False)
        inf = {"d": d, "embeddings": n, "points": 25}
        for streams in (1, fad.SCORE_INF_STREAMS):
            fad.SCORE_INF_STREAMS = streams
            np.random.seed(0)
            fad.score_inf(mu_b, cov_b, emb, steps=3, min_n=500)  # warm-up
            np.random.seed(0)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            res = fad.score_inf(mu_b, cov_b, emb, steps=25, min_n=500)
            inf[f"wall_ms_{streams}_stream"] = (time.perf_counter() - t0) * 1e3
            inf[f"score_{streams}_stream"] = float(res.score)
    print(json.dumps({"what": "dm_frechet_distance device time per solve pair", "fad_inf": inf,
                      "jacobi_graph": os.environ.get("DM_JACOBI_GRAPH", "1") != "0", "rows": rows}, indent=1))


if __name__ == "__main__":
    main()
