"""BASELINE config 5 (FAD part): embedding statistics of 1024 clips sharded over the ranks of one node, ONE all-reduce.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/fad_dist_check.py [--d 768 --frames 499 --clips 1024]

Each rank accumulates the raw moments of its clips (clip i -> rank i mod W) with dm_fad_moments (tcgen05 + TMA), the
packed float64 vector [n | sum x | sum x x^T] is all-reduced once over NCCL, every rank finalises mu / cov.  Rank 0
checks the result against float64 NumPy on the full data and prints one JSON line with device-side timings (max over
ranks)."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffmusic_b200 import fad, parallel  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--d", type=int, default=768)
    ap.add_argument("--frames", type=int, default=499)
    ap.add_argument("--clips", type=int, default=1024)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--check", type=int, default=1)
    a = ap.parse_args()
    rank, world, local = parallel.init_distributed()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    mine = parallel.shard_indices(a.clips, rank, world)
    # synthetic fp16 embeddings per clip, seeded by clip id (so every world size sees the same data)
    blocks = []
    for i in mine:
        g = torch.Generator().manual_seed(7000 + i)
        blocks.append((torch.randn(a.frames, a.d, generator=g) * 0.6 + 0.25).half())
    X = torch.cat(blocks).to(dev)

    def run():
        m = fad.EmbeddingMoments(a.d, device=dev)
        m.update(X)
        m.all_reduce()
        return m, m.finalize()

    run()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(a.iters):
        m, (mu, cov) = run()
    t.record()
    torch.cuda.synchronize()
    ms = torch.tensor([s.elapsed_time(t) / a.iters], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ok = None
    if a.check and rank == 0:
        allX = []
        for i in range(a.clips):
            g = torch.Generator().manual_seed(7000 + i)
            allX.append((torch.randn(a.frames, a.d, generator=g) * 0.6 + 0.25).half())
        A = torch.cat(allX).numpy().astype(np.float64)
        wmu, wcov = A.mean(0), np.cov(A, rowvar=False)
        emu = np.linalg.norm(mu.cpu().numpy() - wmu) / np.linalg.norm(wmu)
        ecov = np.linalg.norm(cov.cpu().numpy() - wcov) / np.linalg.norm(wcov)
        ok = bool(m.count() == A.shape[0] and emu < 1e-6 and ecov < 1e-5)
        print(f"rel err mu {emu:.2e} cov {ecov:.2e}", file=sys.stderr)
    if rank == 0:
        n_total = a.clips * a.frames
        print(json.dumps({"what": "FAD moments + all-reduce + finalize", "n_gpus": world, "clips": a.clips,
                          "frames_per_clip": a.frames, "d": a.d, "ms": float(ms.item()),
                          "embeddings_per_s": n_total / (float(ms.item()) * 1e-3),
                          "tflops_xtx": 2.0 * n_total * a.d * a.d / (float(ms.item()) * 1e-3) / 1e12,
                          "allreduce_bytes": 8 * (1 + a.d + a.d * a.d), "matches_numpy_fp64": ok}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
