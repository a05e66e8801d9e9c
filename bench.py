#!/usr/bin/env python
"""bench.py -- guided denoising steps/s on 10 s clips (BASELINE.json metric), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg1|cfg3|cfg4|cfg5]
                    [--networks stub|hifigan]

A *step* is one pass of the guidance hot path over one batch of synthetic clips: `scheduler.step(...)` of the
host-side mirror = x0 kernel -> (torch stand-ins for vae.decode / vocoder, which stay in PyTorch) -> fused operator +
T_mel + loss + VJP kernels -> torch autograd back through the stand-ins -> fused scheduler-update kernel.
Default workload = BASELINE.json configs[1]: super_resolution (scale 2) + DPS, batch 16 x 10 s @ 16 kHz on one B200.

  value : clip-steps/s with inputs resident in HBM, in the captured graph's static input buffers (CUDA events per
          step, L2 flushed between steps, max over ranks)
  e2e   : the same through the public API (`HostPipelinedStep` over `GraphedGuidedStep`) with PINNED HOST latents in
          and prev_sample + per-clip loss out, every host<->device copy inside a timed bracket; the copies of
          neighbouring steps overlap the step on separate streams (`serial_ms_per_step` = no overlap)
  roofline     : dominant kernel (stft_warp_kernel, the fused STFT / mel / loss / VJP pass), algorithmic bytes / CUDA-event time, vs MEASURED_PEAKS hbm_gbs
  cpu_baseline : the CPU oracle (torch restatement of the reference's scheduler.step) on this box's host cores, on a
                 bounded sample of the same workload
`--impl reference` times that CPU oracle alone, with every host thread, and prints the same line with impl=reference.

`--workload cfg5` (BASELINE configs[4], the only collective on the path): FAD statistics of 1024 clips x 499 frames x 768
fp16 embeddings sharded clip i -> rank i mod N (strong scaling).  A step = clear the accumulator -> dm_fad_moments
(tcgen05) -> exchange (one kernel per rank over NVLink peer memory, csrc/fad_exchange.cu; nothing to do at N = 1) ->
dm_fad_finalize_sym; the JSON line carries the per-phase split, roofline.bound = "tensor" and np.cov (float64) as the CPU
baseline.  `--networks hifigan` swaps the stand-in vocoder for a random-init SpeechT5HifiGan.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from tests import stubs  # noqa: E402  (synthetic inputs + torch stand-ins for the networks that stay in PyTorch)

L10 = 160000
WORKLOADS = {
    # name: (scheduler, operator, batch per GPU, eta, rate, description)
    "cfg1": ("ddim", "inpainting", 1, 0.0, None, "music_inpainting box[2s,3s) + DDIM, 1 x 10 s clip"),
    "cfg2": ("dps", "super_resolution", 16, 0.0, 5e-4, "super_resolution scale 2 + DPS, 16 x 10 s clips per GPU"),
    "cfg3": ("dsg", "phase_retrieval", 8, 1.0, 0.08, "phase_retrieval |STFT| + DSG, 8 x 10 s clips per GPU"),
    "cfg4": ("diffmusic", "dereverberation", 16, 1.0, 0.08, "dereverberation K=5000 + DiffMusic, 16 x 10 s per GPU"),
}


# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel, from the committed
# `ncu --set full` captures (profiles/README.md); null where no capture of that workload exists.
NCU_TRAFFIC_BYTES = {"cfg2": 10.85e6}
NCU_TRAFFIC_SOURCE = "profiles/r02/stft_warp_full_raw.csv (ncu --set full, per launch)"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi SM clock / throttle-reason samples while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_networks(kind, device):
    """(vae, vocoder, description): the networks that stay in PyTorch.  "stub": the deterministic stand-ins of
    tests/stubs.py; "hifigan": the same VAE-decoder stand-in (diffusers' AutoencoderKL is not in the image) with a
    random-init `transformers.SpeechT5HifiGan` of the MusicLDM vocoder shape."""
    vae = stubs.StubVAE().to(device)
    if kind == "hifigan":
        from transformers import SpeechT5HifiGan, SpeechT5HifiGanConfig
        cfg = SpeechT5HifiGanConfig(model_in_dim=64, sampling_rate=16000, upsample_rates=[5, 4, 2, 2, 2],
                                    upsample_kernel_sizes=[16, 16, 8, 4, 4], upsample_initial_channel=1024,
                                    normalize_before=False)
        torch.manual_seed(0)
        voc = SpeechT5HifiGan(cfg).eval().to(device)
        for q in voc.parameters():
            q.requires_grad_(False)
        return vae, voc, ("random-init transformers.SpeechT5HifiGan vocoder (55 M parameters, MusicLDM shape) + torch "
                          "stand-in for vae.decode (stay in PyTorch, inside the step)")
    return vae, stubs.StubVocoder().to(device), ("torch stand-ins for vae.decode / vocoder (stay in PyTorch, inside the "
                                                 "step)")


def config_dict(args, B, mode):
    """the `config` object of the JSON line: identical for the product arm and the reference arm of one workload"""
    return {"workload": WORKLOADS[args.workload][5], "name": args.workload, "clips_per_gpu": B, "mode": mode,
            "l2": "flushed between timed steps (256 MB write)", "networks": make_networks_description(args.networks)}


def make_networks_description(kind):
    if kind == "hifigan":
        return ("random-init transformers.SpeechT5HifiGan vocoder (55 M parameters, MusicLDM shape) + torch stand-in for "
                "vae.decode (stay in PyTorch, inside the step)")
    return "torch stand-ins for vae.decode / vocoder (stay in PyTorch, inside the step)"


def build_case(workload, device, first_clip=0, batch=None, networks="stub"):
    import diffmusic_b200 as dm
    sched_name, op_name, B, eta, rate, _ = WORKLOADS[workload]
    B = batch or B
    noiser = dm.get_noiser("gaussian", 0.0)
    if op_name == "inpainting":
        op = dm.MusicInpaintingOperator(10, 16000, "box", 2, 3, 0.3, 0.1, 1, noiser=noiser)
    elif op_name == "super_resolution":
        op = dm.SuperResolutionOperator(16000, scale=2, noiser=noiser)
    elif op_name == "phase_retrieval":
        op = dm.PhaseRetrievalOperator(1024, 160, 1024, noiser=noiser)
    else:
        op = dm.MusicDereverberationOperator(ir_length=5000, decay_factor=0.99, noiser=noiser)
    sched = dm.get_scheduler(sched_name)(operator=op, **stubs.MUSICLDM_SCHED)
    sched.set_timesteps(500)
    torch.manual_seed(0)
    meas = op.forward(stubs.synth_clips(1, L10, first=50).to(device))
    x, e = stubs.synth_latents(B, 250, first=first_clip)
    vae, voc, _ = make_networks(networks, device)
    kw = dict(eta=eta, measurement=meas, vae=vae, vocoder=voc, original_waveform_length=L10)
    if rate is not None:
        kw.update(ip_guidance_rate=rate, supervised_space="mel_spectrogram")
    return sched, op, x, e, kw, B


def cpu_oracle_rate(workload, clips, steps, warmup, threads, networks="stub"):
    """clip-steps/s of the CPU oracle (restated reference scheduler.step, torch CPU) on `clips` clips per step."""
    from oracle import operators as oo
    from oracle import steps as osteps
    torch.set_num_threads(threads)
    sched_name, op_name, _, eta, rate, _ = WORKLOADS[workload]
    if op_name == "inpainting":
        op = oo.OracleOperator("inpainting", mask=oo.inpaint_mask(10, 16000, "box", 2, 3))
    elif op_name == "super_resolution":
        op = oo.OracleOperator("super_resolution", scale=2)
    elif op_name == "phase_retrieval":
        op = oo.OracleOperator("phase_retrieval")
    else:
        op = oo.OracleOperator("dereverberation", ir_length=5000, decay_factor=0.99)
    base = osteps.make_base(**stubs.MUSICLDM_SCHED)
    base.set_timesteps(500)
    torch.manual_seed(0)
    meas = op.forward(stubs.synth_clips(1, L10, first=50))
    x, e = stubs.synth_latents(clips, 250)
    vae, voc, _ = make_networks(networks, "cpu")
    ts = [999, 501, 1]
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        osteps.per_clip_step(sched_name, base, op, e, ts[i % 3], x, generators=stubs.step_generators(clips),
                             measurement=meas, eta=eta, ip_guidance_rate=rate, vae=vae, vocoder=voc,
                             original_waveform_length=L10, supervised_space="mel_spectrogram")
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return clips / statistics.mean(times), statistics.mean(times)


def run_reference(args, rank, world):
    if rank != 0:
        return
    if args.workload == "cfg5":
        return run_fad_reference(args)
    threads = os.cpu_count() or 1
    # the workload's own batch per step (dereverberation and the HiFi-GAN vocoder: one clip -- seconds of CPU work each)
    B = WORKLOADS[args.workload][2]
    clips = B if (WORKLOADS[args.workload][1] != "dereverberation" and args.networks == "stub") else 1
    rate, sec = cpu_oracle_rate(args.workload, clips, args.steps, args.warmup, threads, args.networks)
    sample = f"{clips} clip(s) per step of the same workload, {args.steps} steps, oracle per_clip_step on CPU"
    line = {"impl": "reference", "metric": "guided denoising steps/s (10 s clips)", "value": rate,
            "unit": "clip-steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": config_dict(args, B, "eager scheduler.step" if args.eager else
                                  "CUDA-graph replay of scheduler.step"),
            "cpu_baseline": {"value": rate, "unit": "clip-steps/s", "cores": threads, "kind": "port",
                             "sample": sample},
            "e2e": {"value": rate, "unit": "clip-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ cfg5: FAD statistics
FAD_CLIPS, FAD_FRAMES, FAD_D = 1024, 499, 768
FAD_METRIC = "FAD embedding statistics (mean + covariance of 1024 clips x 499 x 768 fp16)"


def fad_clip(i):
    g = torch.Generator().manual_seed(7000 + i)
    return (torch.randn(FAD_FRAMES, FAD_D, generator=g) * 0.6 + 0.25).half()


def fad_cpu_rate(clips, threads):
    """embeddings/s of the reference arithmetic (np.mean + np.cov in float64, fadtk/fad.py:41-47) on `clips` clips"""
    import numpy as np
    torch.set_num_threads(threads)
    X = torch.cat([fad_clip(i) for i in range(clips)]).numpy()
    t0 = time.perf_counter()
    np.mean(X, axis=0)
    np.cov(X, rowvar=False)
    sec = time.perf_counter() - t0
    return X.shape[0] / sec, sec


def run_fad_reference(args):
    threads = os.cpu_count() or 1
    clips = 64
    rates = [fad_cpu_rate(clips, threads) for _ in range(max(1, min(args.steps, 5)))]
    rate, sec = max(r for r, _ in rates), min(s for _, s in rates)
    line = {"impl": "reference", "metric": FAD_METRIC, "value": rate, "unit": "embeddings/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": fad_config(),
            "cpu_baseline": {"value": rate, "unit": "embeddings/s", "cores": threads, "kind": "port",
                             "sample": f"np.mean + np.cov (float64) of {clips} of the 1024 clips, best of {len(rates)}"},
            "e2e": {"value": rate, "unit": "embeddings/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def fad_config():
    return {"workload": "FAD moments + exchange + finalize, 1024 clips x 499 frames x d = 768 fp16, clip i -> rank i mod N",
            "name": "cfg5", "clips": FAD_CLIPS, "frames_per_clip": FAD_FRAMES, "d": FAD_D,
            "l2": "flushed between timed steps (256 MB write)"}


def run_fad(args, rank, world, local, device, dist):
    """BASELINE configs[4]: the rank's clips are resident in HBM (`value`) or in pinned host memory (`e2e`); every rank
    ends a step holding the statistics of ALL clips."""
    import faulthandler
    import numpy as np
    from diffmusic_b200 import _lib, fad, parallel
    faulthandler.dump_traceback_later(240, exit=True)  # a stuck exchange must not hold the box
    mine = parallel.shard_indices(FAD_CLIPS, rank, world)
    X_host = torch.cat([fad_clip(i) for i in mine]).pin_memory()
    X = X_host.to(device)
    n_rows = FAD_CLIPS * FAD_FRAMES
    xkind = os.environ.get("DM_FAD_EXCHANGE", "peer")  # "peer" (default), "peer_oneshot", "peer_push" (A/B)
    mom = fad.EmbeddingMoments(FAD_D, device=device, exchange=xkind if world > 1 else None)
    mom_reset = mom.reset
    flush = torch.empty(256 * 1024 * 1024 // 4, device=device, dtype=torch.float32)
    mu_pin = torch.empty(FAD_D, dtype=torch.float64).pin_memory()
    cov_pin = torch.empty(FAD_D, FAD_D, dtype=torch.float64).pin_memory()
    X_stage = torch.empty_like(X)

    def step(host, ev=None):
        src = X
        if host:
            X_stage.copy_(X_host, non_blocking=True)
            src = X_stage
        mom_reset()
        mom.update(src)
        if ev is not None:
            ev[0].record()
        if world > 1:
            mom.all_reduce()
        if ev is not None:
            ev[1].record()
        mu, cov = mom.finalize()
        if host:
            mu_pin.copy_(mu, non_blocking=True)
            cov_pin.copy_(cov, non_blocking=True)
        return mu, cov

    def region(host):
        for _ in range(args.warmup):
            step(host)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        n0 = _lib.launch_count()
        torch.cuda._sleep(int(2e6))  # ~1 ms head start: the host enqueues ahead, the brackets hold device time only
        br = []
        for i in range(args.steps):
            flush.fill_(float(i))
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            evs[0].record()
            step(host, ev=(evs[1], evs[2]))
            evs[3].record()
            br.append(evs)
        torch.cuda.synchronize()
        launches = _lib.launch_count() - n0
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        tot = sum(e[0].elapsed_time(e[3]) for e in br)
        phases = [sum(e[k].elapsed_time(e[k + 1]) for e in br) / len(br) for k in range(3)]
        tt = torch.tensor([tot] + phases, device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return [float(v) for v in tt.tolist()], launches

    clk = ClockSampler(local) if rank == 0 else None
    if clk:
        clk.__enter__()
    (tot, ph_m, ph_x, ph_f), launches = region(False)
    (tot_h, _, _, _), _ = region(True)
    if clk:
        clk.__exit__()
    mu, cov = step(False)
    torch.cuda.synchronize()
    ok = None
    if rank == 0 and args.fad_check:  # parity of the all-reduced statistics against float64 NumPy on ALL clips (~1 min)
        A = torch.cat([fad_clip(i) for i in range(FAD_CLIPS)]).numpy().astype(np.float64)
        wmu, wcov = A.mean(0), np.cov(A, rowvar=False)
        emu = np.linalg.norm(mu.cpu().numpy() - wmu) / np.linalg.norm(wmu)
        ecov = np.linalg.norm(cov.cpu().numpy() - wcov) / np.linalg.norm(wcov)
        ok = bool(mom.count() == n_rows and emu < 1e-6 and ecov < 1e-5)
    if world > 1:
        dist.barrier()
    mom.close()
    faulthandler.cancel_dump_traceback_later()
    if rank != 0:
        return
    ms = tot / args.steps
    value = n_rows / (ms * 1e-3)
    e2e = n_rows / (tot_h / args.steps * 1e-3)
    peaks_json = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    peak = float(peaks_json.get("bf16_tflops", 2250.0))
    nblk = (FAD_D + 127) // 128
    executed = 2.0 * (n_rows / world) * 128 * 128 * (nblk * (nblk + 1) // 2)   # upper-triangle tiles of this rank
    achieved = executed / (ph_m * 1e-3) / 1e12
    line = {"metric": FAD_METRIC, "value": value, "unit": "embeddings/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f16 x f16 -> f32 (tcgen05), f64 accumulation", "data": "synthetic",
            "config": fad_config(),
            "phases_ms": {"moments": ph_m, "exchange": ph_x, "finalize": ph_f,
                          "exchange_kind": (("one kernel per rank over NVLink peer memory (CUDA IPC), upper triangle only: "
                                             + ("every rank reads every peer's triangle (one-shot)"
                                                if (xkind == "peer_oneshot" or (xkind == "peer" and world <= 2)) else
                                                "rank r reduces rows r, r+W, ... and pushes them to every rank"))
                                            if world > 1 else "none (one rank)"),
                          "packed_bytes": 8 * int(_lib.load().dm_fad_packed_doubles(FAD_D))},
            "matches_numpy_fp64": ok,
            "e2e": {"value": e2e, "unit": "embeddings/s", "h2d_bytes_per_step": X_host.numel() * 2,
                    "d2h_bytes_per_step": 8 * (FAD_D + FAD_D * FAD_D), "ms_per_step": tot_h / args.steps},
            "gpu_launches": launches,
            "roofline": {"bound": "tensor", "kernel": "fad_xtx_tc_kernel", "achieved": achieved, "peak": peak,
                         "unit": "TFLOP/s", "frac": achieved / peak, "traffic": None,
                         "flops_executed_per_launch": executed, "ms_per_launch": ph_m,
                         "peak_source": "measured (MEASURED_PEAKS.json bf16_tflops)" if peaks_json else "nominal",
                         "note": "executed = the upper-triangle 128 x 128 tiles of X^T X (58 % of 2 N d^2 at d = 768); "
                                 "ms_per_launch = the moments phase (clear + tcgen05 kernels) of one step"},
            "clocks": clk.summary() if clk else None}
    if not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        rate, sec = fad_cpu_rate(64, threads)
        line["cpu_baseline"] = {"value": rate, "unit": "embeddings/s", "cores": threads, "kind": "port",
                                "sample": f"np.mean + np.cov (float64, fadtk/fad.py:41-47) of 64 of the 1024 clips, "
                                          f"{sec:.2f} s"}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS) + ["cfg5"])
    ap.add_argument("--networks", default="stub", choices=["stub", "hifigan"],
                    help="vocoder inside the step: torch stand-in (default) or random-init SpeechT5HifiGan")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--fad-check", action="store_true",
                    help="cfg5: compare the all-reduced statistics with float64 NumPy on all 1024 clips (about a minute)")
    ap.add_argument("--serial-e2e", action="store_true", help="e2e with the copies on the compute stream (no overlap)")
    ap.add_argument("--eager", action="store_true", help="call scheduler.step eagerly instead of replaying a CUDA graph")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch.distributed as dist
    from diffmusic_b200 import _lib
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:  # every rank enqueues from its own cores (max-over-ranks timing is sensitive to host interference)
        try:
            cores = sorted(os.sched_getaffinity(0))
            share = max(1, len(cores) // world)
            os.sched_setaffinity(0, cores[local * share:(local + 1) * share] or cores)
            torch.set_num_threads(max(1, min(share, 4)))
        except (AttributeError, OSError):
            pass
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner to STDOUT when the first communicator comes up; stdout must carry exactly one
        # JSON line, so fd 1 points at stderr until the communicator exists
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=device)
            warm = torch.zeros(1, device=device)
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    if args.workload == "cfg5":
        run_fad(args, rank, world, local, device, dist)
        if world > 1:
            dist.destroy_process_group()
        return

    sched, op, x_h, e_h, kw, B = build_case(args.workload, device, first_clip=rank * 64, networks=args.networks)
    x_d, e_d = x_h.to(device), e_h.to(device)
    x_pin, e_pin = x_h.pin_memory(), e_h.pin_memory()
    prev_pin = torch.empty_like(x_h).pin_memory()
    loss_pin = torch.empty(B).pin_memory()
    flush = torch.empty(256 * 1024 * 1024 // 4, device=device, dtype=torch.float32)  # 256 MB > 126 MB L2
    ts = [int(t) for t in sched.timesteps[: args.steps + args.warmup]]
    gens = stubs.step_generators(B, first=rank * 64, device=device) if kw["eta"] > 0 else None

    # event pair around the dominant kernel (dm_stft_guidance), on the launching stream
    dom_events = []
    dom_names = set()
    orig_call = _lib.call

    other_events = {}

    def timed_call(name, *a):
        if timed_call.on:
            s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            orig_call(name, *a)
            t.record()
            if name in ("dm_stft_guidance", "dm_stft_guidance_io", "dm_stft_guidance_fir2"):
                dom_events.append((s, t))
                dom_names.add(name)
            else:
                other_events.setdefault(name, []).append((s, t))
        else:
            orig_call(name, *a)

    timed_call.on = False
    _lib.call = timed_call
    import diffmusic_b200.operators as _ops_mod
    import diffmusic_b200.schedulers as _sch_mod
    _ops_mod._lib.call = timed_call
    _sch_mod._lib.call = timed_call

    import diffmusic_b200 as dm
    graphed = None if args.eager else dm.GraphedGuidedStep(sched, tuple(x_d.shape), clone_outputs=False, **kw)
    # second capture of the same step: with two graphs HostPipelinedStep uploads straight into the static inputs of the
    # graph that consumes them (no device-to-device staging)
    graphed_b = None if (args.eager or args.serial_e2e) else dm.GraphedGuidedStep(sched, tuple(x_d.shape),
                                                                                  clone_outputs=False, **kw)

    def one_step(i, host, eager=False):
        t = ts[i % len(ts)]
        xs, es = (x_pin, e_pin) if host else (x_d, e_d)
        if graphed is None or eager:
            if host:
                xs, es = xs.to(device, non_blocking=True), es.to(device, non_blocking=True)
            out = sched.step(es, t, xs, generator=gens, **kw)
        elif host:
            out = graphed(es, t, xs, generator=gens)  # H2D copies into its static buffers, then replays
        else:
            out = graphed.replay_in_place(t, generator=gens)  # inputs are resident in the graph's static buffers
        if host:
            prev_pin.copy_(out.prev_sample, non_blocking=True)
            lo = out.loss_per_clip if out.loss_per_clip is not None else out.loss.float().to(device)
            loss_pin[: lo.numel()].copy_(lo.reshape(-1), non_blocking=True)
        return out

    host_ms = {}

    def pipelined_region():
        """e2e through HostPipelinedStep: pinned-host latents in, prev_sample + per-clip loss out, uploads of step i+1
        and downloads of step i-1 overlapped with step i.  Every copy is issued after the start event of a timed
        bracket and waited for before the end event of one (the last download gets a bracket of its own)."""
        xs_pin = [x_pin, x_h.clone().pin_memory()]
        es_pin = [e_pin, e_h.clone().pin_memory()]
        prevs_pin = [prev_pin, torch.empty_like(x_h).pin_memory()]
        losses_pin = [loss_pin, torch.empty(B).pin_memory()]

        def sequence(n, first, timed):
            pipe = dm.HostPipelinedStep(graphed, graphed_b)
            brackets = []
            if timed:  # same gate as the device-resident loop: the host enqueues ahead of the device
                torch.cuda._sleep(int(max(4e6, 3e5 * n)))
            t_host0 = time.perf_counter()
            for i in range(n):
                if timed:
                    flush.fill_(float(i))
                s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                if i == 0:
                    pipe.prefetch(es_pin[0], xs_pin[0], after=s)
                # step first, then the next upload: the step's own small host->device copies (coefficients, the
                # dereverberation impulse response) must not queue behind 4 MB of latents on the copy engine
                pipe.step(ts[(first + i) % len(ts)], prevs_pin[i % 2], losses_pin[i % 2], generator=gens)
                if i + 1 < n:
                    pipe.prefetch(es_pin[(i + 1) % 2], xs_pin[(i + 1) % 2], after=s)
                t.record()
                brackets.append((s, t))
            host_ms["e2e"] = (time.perf_counter() - t_host0) * 1e3 / max(n, 1)  # host enqueue time per step
            s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            pipe.drain()
            t.record()
            brackets.append((s, t))
            return brackets

        sequence(args.warmup, 0, False)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        brackets = sequence(args.steps, args.warmup, True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if not torch.isfinite(prevs_pin[(args.steps - 1) % 2]).all():
            raise RuntimeError("pipelined e2e produced non-finite latents")
        total_ms = sum(s.elapsed_time(t) for s, t in brackets)
        tt = torch.tensor([total_ms], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    def timed_region(host):
        per_step = []
        if graphed is not None and not host:  # "inputs already resident in HBM": the step reads them where they are
            graphed.x.copy_(x_d)
            graphed.e.copy_(e_d)
        for i in range(args.warmup):
            one_step(i, host)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        launches0 = _lib.launch_count()
        # Gate: the device first spins for ~20 us per timed step (at least 2 ms) while the host enqueues the loop, so no
        # event bracket below contains host enqueue time or scheduling jitter -- with N ranks on one host the maximum
        # over ranks would otherwise measure the slowest Python thread, not the GPUs.
        if not host:
            torch.cuda._sleep(int(max(4e6, 4e4 * args.steps * (8 if graphed is None else 1))))
        t_host0 = time.perf_counter()
        for i in range(args.steps):
            flush.fill_(float(i))  # evict L2 between timed iterations (not timed)
            s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            one_step(args.warmup + i, host)
            t.record()
            per_step.append((s, t))
        host_ms["host" if host else "device"] = (time.perf_counter() - t_host0) * 1e3 / args.steps
        torch.cuda.synchronize()
        launches = _lib.launch_count() - launches0
        if graphed is not None:
            launches += graphed.kernels_per_replay * args.steps  # kernel nodes replayed from the captured graph
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        total_ms = sum(s.elapsed_time(t) for s, t in per_step)
        tt = torch.tensor([total_ms], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item()), launches

    import contextlib
    with (ClockSampler(local) if rank == 0 else contextlib.nullcontext()) as clk:  # one sampler process per node
        total_ms, launches = timed_region(host=False)
        e2e_serial_ms, _ = timed_region(host=True)
        e2e_ms = pipelined_region() if (graphed is not None and not args.serial_e2e) else e2e_serial_ms
        # the dominant kernel, timed live with CUDA events on its launching stream inside eager steps (a graph replay
        # has no per-kernel event hooks), L2 flushed before every step as above
        # An eager step is host-bound (~0.9 ms of Python dispatch for ~0.2 ms of device work): with an empty queue every
        # event bracket would also contain that kernel's launch latency.  ~1.2 ms of L2-flush fills are queued first,
        # so the host runs ahead and the step's kernels execute back to back, as they do inside the replayed graph.
        timed_call.on = True
        for i in range(args.steps):
            for j in range(24):
                flush.fill_(float(i + j))
            one_step(args.warmup + i, False, eager=True)
        torch.cuda.synchronize()
        timed_call.on = False
        dom_ms = [s.elapsed_time(t) for s, t in dom_events]
    clocks = clk.summary() if clk is not None else None

    if rank == 0:
        ms_step = total_ms / args.steps
        value = world * B / (ms_step * 1e-3)
        e2e_value = world * B / (e2e_ms / args.steps * 1e-3)
        # algorithmic bytes of the dominant kernel per launch: signal in + reference mel in + padded
        # cotangent out, per clip, times the clips of one launch
        op_name = WORKLOADS[args.workload][1]
        Ly = {"super_resolution": L10 // 2, "dereverberation": L10 + 1}.get(op_name, L10)
        T = 1 + Ly // 160
        rows = 64
        fused_fir = "dm_stft_guidance_fir2" in dom_names  # cfg2: the kernel reads the 16 kHz waveform and resamples it itself
        bytes_launch = B * (4 * (L10 if fused_fir else Ly) + 4 * rows * T + 4 * (Ly + 1024))
        peak, peak_src = peaks()
        dom = statistics.mean(dom_ms) if dom_ms else None
        achieved = bytes_launch / (dom * 1e-3) / 1e9 if dom else None
        roofline = {"bound": "hbm", "kernel": "stft_warp_kernel", "achieved": achieved, "peak": peak,
                    "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                    "traffic": NCU_TRAFFIC_BYTES.get(args.workload) if dom else None,
                    "traffic_source": NCU_TRAFFIC_SOURCE,
                    "bytes_per_launch": bytes_launch, "ms_per_launch": dom, "launches_timed": len(dom_ms),
                    "peak_source": peak_src,
                    "note": "compute/shared-memory bound FFT kernel: algorithmic HBM bytes are the floor, see DESIGN.md"
                            + ("; scale-2 resampling fused into the kernel (dm_stft_guidance_fir2): signal in = the 16 kHz "
                               "waveform, the 8 kHz signal never exists in HBM" if fused_fir else "")}
        # the other kernels of ours inside the step, same live timing (eager steps, cold L2): algorithmic bytes per call
        lat = 4 * x_h.numel()
        Lw = L10 + 32  # stand-in vocoder output length
        alg = {"dm_sched_x0_io": 4 * lat, "dm_sched_ddim_update_io": 3 * lat, "dm_sched_dps_update_io": 4 * lat,
               "dm_sched_mpgd_update_io": 5 * lat, "dm_sched_dsg_update_io": 5 * lat,
               "dm_sched_diffmusic_update_io": 5 * lat, "dm_resample_fwd_io": B * (4 * L10 + 4 * Ly),
               "dm_resample_adjoint_io": B * (4 * (Ly + 1024) + 4 * L10),
               "dm_fold_adjoint_io": B * (4 * (Ly + 1024) + 4 * L10),
               "dm_rir_correlate": B * (4 * L10 + 4 * Ly), "dm_rir_adjoint": B * (4 * (Ly + 1024) + 4 * L10)}
        others = []
        for name, evs in sorted(other_events.items()):
            ms = statistics.mean(s.elapsed_time(t) for s, t in evs)
            row = {"call": name, "ms": ms, "calls_timed": len(evs)}
            if name in alg and ms > 0:
                row.update(bytes=alg[name], GBps=alg[name] / (ms * 1e-3) / 1e9, frac=alg[name] / (ms * 1e-3) / 1e9 / peak)
            others.append(row)
        roofline["other_calls"] = others
        line = {"metric": "guided denoising steps/s (10 s clips)", "value": value, "unit": "clip-steps/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": config_dict(args, B, "eager scheduler.step" if graphed is None else
                                      "CUDA-graph replay of scheduler.step"),
                "e2e": {"value": e2e_value, "unit": "clip-steps/s", "h2d_bytes_per_step": 2 * x_h.numel() * 4,
                        "d2h_bytes_per_step": x_h.numel() * 4 + B * 4, "ms_per_step": e2e_ms / args.steps,
                        "mode": "serial copies" if e2e_ms is e2e_serial_ms else
                        "HostPipelinedStep over two captured graphs: uploads / downloads of neighbouring steps overlap the step (3 streams), no staging copies",
                        "serial_ms_per_step": e2e_serial_ms / args.steps,
                        "host_enqueue_ms_per_step": host_ms.get("e2e")},
                "host_enqueue_ms_per_step": host_ms.get("device"),
                "gpu_launches": launches, "roofline": roofline, "clocks": clocks}
        if not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            clips = B if (op_name != "dereverberation" and args.networks == "stub") else 1
            _, probe = cpu_oracle_rate(args.workload, clips, 1, 1, threads, args.networks)
            n_cpu = max(3, min(300, int(12.0 / max(probe, 1e-3))))  # ~12 s of host work
            rate, sec = cpu_oracle_rate(args.workload, clips, n_cpu, 0, threads, args.networks)
            line["cpu_baseline"] = {"value": rate, "unit": "clip-steps/s", "cores": threads, "kind": "port",
                                    "sample": f"{clips} clip(s) x {n_cpu} steps of the same workload through the CPU "
                                              f"oracle (oracle/steps.py per_clip_step), {sec:.3f} s per step"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
