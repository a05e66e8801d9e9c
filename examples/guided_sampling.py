"""Guided sampling loop with the drop-in schedulers/operators -- the structure of the reference's pipeline loop
(pipeline_musicldm.py:677-763) with small deterministic stand-ins for the UNet / VAE / vocoder (tests/stubs.py), so it
runs without checkpoints:

    python examples/guided_sampling.py --task super_resolution --scheduler dps --steps 50 --batch 4 [--graph]

What it shows: `diffmusic.schedulers` / `diffmusic.inverse_problem` resolve to diffmusic_b200 through the namespace
drop-in, the pipeline-side calls (`set_timesteps`, `.step(...)`, the NaN restart guard on `out.loss`, `out.prev_sample`)
are the reference's, a batch of clips runs as independent trajectories with per-clip generators, and the whole step can
be replayed as one CUDA graph.
"""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "diffmusic_b200", "dropin"))

from diffmusic.inverse_problem import get_noiser  # noqa: E402  (-> diffmusic_b200.noise)
from diffmusic.inverse_problem.operator import (MusicDereverberationOperator, MusicInpaintingOperator,  # noqa: E402
                                                PhaseRetrievalOperator, SuperResolutionOperator)
from diffmusic.schedulers import get_scheduler  # noqa: E402  (-> diffmusic_b200.schedulers)
from diffmusic_b200 import GraphedGuidedStep  # noqa: E402
from tests import stubs  # noqa: E402

RATES = {"dps": (0.0, 5e-4), "mpgd": (0.0, 0.005), "dsg": (1.0, 0.08), "diffmusic": (1.0, 0.08)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--task", default="super_resolution",
                    choices=["music_inpainting", "super_resolution", "phase_retrieval", "music_dereverberation"])
    ap.add_argument("--scheduler", default="dps", choices=sorted(RATES))
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--seconds", type=int, default=10)
    ap.add_argument("--graph", action="store_true")
    ap.add_argument("--driver", action="store_true", help="run the loop through BatchedGuidedSampler (per-clip NaN restarts)")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    sr, L = 16000, a.seconds * 16000
    noiser = get_noiser("gaussian", 0.0)
    op = {"music_inpainting": lambda: MusicInpaintingOperator(a.seconds, sr, "box", 0.2 * a.seconds, 0.3 * a.seconds, 0.3,
                                                              0.1, 1, noiser=noiser),
          "super_resolution": lambda: SuperResolutionOperator(sample_rate=sr, scale=2, noiser=noiser),
          "phase_retrieval": lambda: PhaseRetrievalOperator(1024, 160, 1024, noiser=noiser),
          "music_dereverberation": lambda: MusicDereverberationOperator(ir_length=5000, decay_factor=0.99,
                                                                        noiser=noiser)}[a.task]()
    sched = get_scheduler(a.scheduler)(operator=op, **stubs.MUSICLDM_SCHED)
    sched.set_timesteps(a.steps)
    vae, vocoder = stubs.StubVAE().to(dev), stubs.StubVocoder().to(dev)
    unet = torch.nn.Conv2d(8, 8, 3, padding=1).to(dev)  # stand-in noise predictor
    measurement = op.forward(stubs.synth_clips(1, L, first=50)).to(dev)  # CPU in -> CPU out, like run.py:286
    generators = [torch.Generator(device=dev).manual_seed(i) for i in range(a.batch)]
    latents = torch.cat([torch.randn(1, 8, a.seconds * 25, 16, generator=g, device=dev) for g in generators])
    latents = latents * sched.init_noise_sigma
    eta, rate = RATES[a.scheduler]
    kw = dict(eta=eta, measurement=measurement, vae=vae, vocoder=vocoder, original_waveform_length=L,
              ip_guidance_rate=rate, supervised_space="mel_spectrogram")
    if a.driver:  # the same loop, owned by the batched-clip driver (diffmusic_b200/driver.py)
        from diffmusic_b200 import BatchedGuidedSampler
        drv = BatchedGuidedSampler(sched, lambda x, t: unet(x), vae, vocoder, num_inference_steps=a.steps,
                                   original_waveform_length=L, latent_shape=(8, a.seconds * 25, 16), eta=eta,
                                   ip_guidance_rate=rate, graph=a.graph)
        drv(measurement, generators)  # first call: lazy initialisation, cuDNN heuristics, graph capture
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = drv(measurement, generators, decode=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"{a.task} / {a.scheduler} through BatchedGuidedSampler: {a.batch} clips x {a.steps} steps in {dt:.3f} s "
              f"({a.batch * a.steps / dt:.0f} clip-steps/s, {'graph' if a.graph else 'eager'}, second call), final per-clip loss "
              f"{[round(float(v), 3) for v in res.loss]}, restarts {res.restarts}, audio {tuple(res.audios.shape)}")
        return
    step = GraphedGuidedStep(sched, tuple(latents.shape), **kw) if a.graph else None
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    restarts = 0
    for t in sched.timesteps:
        with torch.no_grad():
            noise_pred = unet(sched.scale_model_input(latents, t))
        out = step(noise_pred, t, latents, generator=generators) if step is not None else \
            sched.step(noise_pred, t, latents, generator=generators, **kw)
        if torch.isnan(out.loss):  # the reference's restart guard (pipeline_musicldm.py:742-756)
            restarts += 1
            latents = torch.randn_like(latents) * sched.init_noise_sigma
            continue
        latents = out.prev_sample.detach()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{a.task} / {a.scheduler}: {a.batch} clips x {a.steps} steps in {dt:.3f} s "
          f"({a.batch * a.steps / dt:.0f} clip-steps/s, {'graph' if a.graph else 'eager'}), final per-clip loss "
          f"{[round(float(v), 3) for v in out.loss_per_clip]}, restarts {restarts}")


if __name__ == "__main__":
    main()
