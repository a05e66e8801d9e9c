"""Sharded run of the batched-clip driver on the GPUs of one node: clip i -> rank i mod W, no
collective on the guided path; only the per-clip loss table is gathered (logging).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29517 \
        tools/driver_dist_check.py [--clips 32 --steps 20 --scheduler dsg --task music_inpainting --graph]

Every rank builds the same seeded clips / generators and drives its own shard through BatchedGuidedSampler; clip 5 is
poisoned (NaN noise prediction) at the third step of its first attempt, so one rank exercises the per-clip restart.
Rank 0 then re-runs ALL clips alone as one batch and checks that every clip's final distance and restart count from the
sharded run equal the single-process ones; it prints one JSON line (device time of the sharded run, max over ranks)."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import diffmusic_b200 as dm  # noqa: E402
from diffmusic_b200 import parallel  # noqa: E402
from tests import stubs  # noqa: E402

RATES = {"dps": (0.0, 5e-4), "mpgd": (0.0, 0.005), "dsg": (1.0, 0.08), "diffmusic": (1.0, 0.08)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=32)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--seconds", type=int, default=10)
    ap.add_argument("--scheduler", default="dsg", choices=sorted(RATES))
    ap.add_argument("--graph", action="store_true")
    a = ap.parse_args()
    rank, world, local = parallel.init_distributed()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    L, H = a.seconds * 16000, a.seconds * 25
    nz = dm.get_noiser("gaussian", 0.0)
    op = dm.MusicInpaintingOperator(a.seconds, 16000, "box", 0.2 * a.seconds, 0.3 * a.seconds, 0.3, 0.1, 1, noiser=nz)
    sched = dm.get_scheduler(a.scheduler)(operator=op, **stubs.MUSICLDM_SCHED)
    vae, voc = stubs.StubVAE().to(dev), stubs.StubVocoder().to(dev)
    torch.manual_seed(0)
    net = torch.nn.Conv2d(8, 8, 3, padding=1).to(dev)
    eta, rate = RATES[a.scheduler]

    def predictor():
        attempt = {}

        def predict(x, t, clips):
            eps = net(x)
            for row, j in enumerate(clips):
                if int(t) == int(sched.timesteps[0]):
                    attempt[j] = attempt.get(j, 0) + 1
                if j == 5 and int(t) == int(sched.timesteps[2]) and attempt[j] == 1:
                    eps[row] = float("nan")
            return eps
        return predict

    def sampler():
        return dm.BatchedGuidedSampler(sched, predictor(), vae, voc, num_inference_steps=a.steps,
                                       original_waveform_length=L, latent_shape=(8, H, 16), eta=eta,
                                       ip_guidance_rate=rate, graph=a.graph)

    def inputs():
        meas = torch.cat([op.forward(stubs.synth_clips(1, L, first=200 + j).to(dev)) for j in range(a.clips)])
        gens = [torch.Generator(device=dev).manual_seed(900 + j) for j in range(a.clips)]
        return meas, gens

    meas, gens = inputs()
    drv = sampler()
    parallel.run_sharded(drv, meas, gens)  # warm-up: lazy init, graph capture
    meas, gens = inputs()
    drv.noise_predictor = predictor()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    ids, out, table = parallel.run_sharded(drv, meas, gens, gather=True)
    t.record()
    torch.cuda.synchronize()
    ms = torch.tensor([s.elapsed_time(t)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        meas, gens = inputs()
        whole = sampler()(meas, gens)
        worst = max(abs(table[i][0] - float(whole.loss[i])) / abs(float(whole.loss[i])) for i in range(a.clips))
        same = [table[i][1] for i in range(a.clips)] == whole.restarts
        steps_run = a.steps * (a.clips + sum(whole.restarts))
        print(json.dumps({"what": "BatchedGuidedSampler sharded over ranks", "n_gpus": world, "clips": a.clips,
                          "steps": a.steps, "scheduler": a.scheduler, "graph": a.graph, "ms": float(ms.item()),
                          "clip_steps_per_s": steps_run / (float(ms.item()) * 1e-3),
                          "restarts": whole.restarts, "restarts_match_single_process": bool(same),
                          "max_rel_loss_diff_vs_single_process": worst, "ok": bool(same and worst < 1e-4)}),
              flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
