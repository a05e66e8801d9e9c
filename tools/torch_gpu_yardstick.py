"""What the reference's own torch / torchaudio calls cost on the same B200 (eager PyTorch: cuFFT, cuBLAS, cuDNN), next to
the fused kernels that replace them.  Uses the library calls of diffmusic/inverse_problem/operator.py directly
(MelSpectrogram + AmplitudeToDB + clamp, Resample, F.conv1d) with torch autograd for the VJP -- no code of this repo's
oracle and none of the reference's files.

    python tools/torch_gpu_yardstick.py [--batch 16] > gpurun_out/torch_gpu_yardstick.json

Two timings per chain: `*_eager_us` (CUDA events around the Python call: includes host dispatch, which dominates the
torch column) and `*_graph_us` (the same call captured once in a CUDA graph and replayed: DEVICE time of the kernels
only -- the kernel-vs-kernel comparison SURVEY.md section 0.1 asks for).  Also records the TF32 matmul peak of this GPU
(8192^3, allow_tf32) next to the bf16 figure of MEASURED_PEAKS.json: the yardstick of any tensor-core DFT experiment.
"""
import argparse
import json
import os
import statistics
import sys

import torch
import torchaudio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import diffmusic_b200 as dm  # noqa: E402

L = 160000


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(iters):
        s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        t.record()
        torch.cuda.synchronize()
        ms.append(s.elapsed_time(t))
    return statistics.median(ms) * 1e3  # us


def graph_timed(fn, iters=20):
    """device time of `fn` (us): captured once in a CUDA graph, replayed under CUDA events"""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        keep = fn()  # noqa: F841  (outputs stay alive with the graph)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    ms = []
    for _ in range(iters):
        s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        g.replay()
        t.record()
        torch.cuda.synchronize()
        ms.append(s.elapsed_time(t))
    return statistics.median(ms) * 1e3


def tf32_peak():
    n = 8192
    a, b = torch.randn(n, n, device="cuda"), torch.randn(n, n, device="cuda")
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        for _ in range(3):
            a @ b
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(10):
            s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            a @ b
            t.record()
            torch.cuda.synchronize()
            best = min(best, s.elapsed_time(t))
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    B = ap.parse_args().batch
    dev = torch.device("cuda", 0)
    wav = (0.1 * torch.randn(B, L, device=dev)).requires_grad_(True)
    ref = 0.1 * torch.randn(1, L, device=dev)
    wav2mel = torchaudio.transforms.MelSpectrogram(sample_rate=16000, n_fft=1024, win_length=1024, hop_length=160,
                                                   n_mels=64, f_min=0.0, f_max=8000.0, power=2.0).to(dev)
    todb = torchaudio.transforms.AmplitudeToDB(stype="power", top_db=None).to(dev)
    resample = torchaudio.transforms.Resample(16000, 8000).to(dev)

    def t_mel(x):
        return torch.clamp(todb(wav2mel(x)), -80, 80)

    rows = []

    def add(name, torch_fn, ours_fn):
        row = {"chain": name, "torch_eager_us": timed(torch_fn), "fused_kernels_us": timed(ours_fn)}
        for key, fn in (("torch_graph_us", torch_fn), ("fused_kernels_graph_us", ours_fn)):
            try:
                row[key] = graph_timed(fn)
            except Exception as exc:  # a library call that cannot be captured: keep the eager number
                import traceback
                row[key + "_error"] = traceback.format_exc()[-600:]
                torch.cuda.synchronize()
        rows.append(row)

    nz = dm.get_noiser("gaussian", 0.0)
    # identity operator: loss + gradient in mel space
    ident = dm.IdentityOperator(16000)
    ref_mel = t_mel(ref).detach()

    def torch_identity():
        loss = torch.linalg.norm(ref_mel - t_mel(wav))
        torch.autograd.grad(loss, wav)

    add("T_mel + loss + VJP (identity / inpainting chain)", torch_identity,
        lambda: ident.fused_loss_and_grad(wav.detach(), ref, "mel_spectrogram"))
    # super-resolution
    sr = dm.SuperResolutionOperator(16000, scale=2, noiser=nz)
    ref_lo_mel = t_mel(resample(ref)).detach()
    meas_sr = sr.forward(ref)

    def torch_sr():
        loss = torch.linalg.norm(ref_lo_mel - t_mel(resample(wav)))
        torch.autograd.grad(loss, wav)

    add("Resample/2 + T_mel + loss + VJP (super-resolution chain, BASELINE config 2)", torch_sr,
        lambda: sr.fused_loss_and_grad(wav.detach(), meas_sr, "mel_spectrogram"))
    # dereverberation
    K = 5000
    ir = torch.cumsum(torch.randn(1, K), 1) * 0.99
    ir = (ir / ir.abs().max()).to(dev)
    dv = dm.MusicDereverberationOperator(ir_length=K, decay_factor=0.99, noiser=nz)
    torch.manual_seed(0)
    meas_dv = dv.forward(ref)
    ref_dv_mel = t_mel(meas_dv).detach()

    def torch_dv():
        y = torch.nn.functional.conv1d(wav[:, None, :], ir[:, None, :], padding=K // 2)[:, 0, :]
        loss = torch.linalg.norm(ref_dv_mel - t_mel(y))
        torch.autograd.grad(loss, wav)

    add("conv1d K=5000 + T_mel + loss + VJP (dereverberation chain, BASELINE config 4)", torch_dv,
        lambda: dv.fused_loss_and_grad(wav.detach(), meas_dv, "mel_spectrogram"))
    for r in rows:
        r["speedup"] = r["torch_eager_us"] / r["fused_kernels_us"]
        if "torch_graph_us" in r and "fused_kernels_graph_us" in r:
            r["device_time_speedup"] = r["torch_graph_us"] / r["fused_kernels_graph_us"]
    print(json.dumps({"what": "reference's torch/torchaudio calls (eager, same GPU) vs the fused kernels", "batch": B,
                      "tf32_matmul_8192_tflops": tf32_peak(),
                      "clip_samples": L, "note": "whole-batch norm in the torch column (as the reference writes it), "
                      "per-clip norms in ours; both include Python dispatch, no L2 flush", "rows": rows}, indent=1))


if __name__ == "__main__":
    main()
