"""Operator-level bandwidth table: "operator HBM GB/s vs peak" of BASELINE.json, no networks in the loop.

    python tools/operator_bench.py [--batch 64] [--iters 20] > gpurun_out/operator_bench.json

For every operator of the guidance path: A(x) forward, transform, and the fused loss + VJP chain (what scheduler.step
calls), plus the scheduler update kernels, on B x 10 s clips (16 kHz) resident in HBM.  Timing: CUDA events around each
call, L2 flushed (256 MB write) before every timed call, median of `iters`.  GB/s = ALGORITHMIC bytes (SURVEY.md 8d:
signal in + cached reference in + result out; intermediates that a fused chain keeps on chip do not count) / time;
`frac` is against MEASURED_PEAKS.json hbm_gbs (fallback 6650).  The FFT-bearing chains are shared-memory / issue bound,
not HBM bound -- their fraction is reported for completeness.
"""
import argparse
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import diffmusic_b200 as dm  # noqa: E402
from tests import stubs  # noqa: E402

L = 160000


def peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return float(json.load(open(p))["hbm_gbs"]) if os.path.exists(p) else 6650.0


def timed(fn, iters, flush):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ms = []
    for i in range(iters):
        flush.fill_(float(i))
        s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        t.record()
        torch.cuda.synchronize()
        ms.append(s.elapsed_time(t))
    return statistics.median(ms)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--torch-yardstick", action="store_true")
    a = ap.parse_args()
    B, dev = a.batch, torch.device("cuda", 0)
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    pk = peak()
    wav = (0.1 * torch.randn(B, L, device=dev)).contiguous()
    ref_wav = stubs.synth_clips(1, L, first=50).to(dev)
    nz = dm.get_noiser("gaussian", 0.0)
    ops = {
        "inpainting": dm.MusicInpaintingOperator(10, 16000, "box", 2, 3, 0.3, 0.1, 1, noiser=nz),
        "super_resolution": dm.SuperResolutionOperator(16000, scale=2, noiser=nz),
        "dereverberation": dm.MusicDereverberationOperator(ir_length=5000, decay_factor=0.99, noiser=nz),
        "phase_retrieval": dm.PhaseRetrievalOperator(1024, 160, 1024, noiser=nz),
    }
    rows = []

    def add(name, fn, nbytes, note=""):
        ms = timed(fn, a.iters, flush)
        gbs = nbytes / (ms * 1e-3) / 1e9
        rows.append({"op": name, "ms": ms, "algorithmic_MB": nbytes / 1e6, "GBps": gbs, "frac_of_hbm_peak": gbs / pk,
                     "note": note})

    F, T, M = 513, 1001, 64
    for name, op in ops.items():
        torch.manual_seed(0)
        meas = op.forward(ref_wav)
        Ly = {"super_resolution": L // 2, "dereverberation": L + 1}.get(name, L)
        Ty = 1 + Ly // 160
        out_bytes = 4 * F * T if name == "phase_retrieval" else 4 * Ly
        extra = 4 * 5000 if name == "dereverberation" else 0
        add(f"{name}.forward", lambda op=op: op.forward(wav), B * (4 * L + out_bytes + extra))
        y = op.forward(wav)
        tr_in = 4 * F * T if name == "phase_retrieval" else 4 * Ly
        add(f"{name}.transform", lambda op=op, y=y: op.transform(y), B * (tr_in + 4 * M * Ty),
            "FFT-bound" if name != "phase_retrieval" else "")
        op.fused_loss_and_grad(wav, meas, "mel_spectrogram")  # caches transform(measurement)
        add(f"{name}.loss+vjp[mel]", lambda op=op, meas=meas: op.fused_loss_and_grad(wav, meas, "mel_spectrogram"),
            B * (4 * L + 4 * L + extra) + 4 * M * Ty, "fused chain, FFT-bound; ref mel shared by the batch")
        wav_space = 12 * L if name == "inpainting" else 4 * L + 4 * L + (4 * F * T if name == "phase_retrieval"
                                                                       else 4 * Ly) / B + extra
        add(f"{name}.loss+vjp[wav]", lambda op=op, meas=meas: op.fused_loss_and_grad(wav, meas, "wav_form"),
            B * wav_space if name == "inpainting" else B * (8 * L + extra) + (4 * F * T if name == "phase_retrieval"
                                                                               else 4 * Ly),
            "FFT-bound" if name in ("phase_retrieval", "dereverberation") else "")
    # scheduler kernels on (B, 8, 250, 16) latents
    x, e = (torch.randn(B, 8, 250, 16, device=dev) for _ in range(2))
    g = torch.randn_like(x) * 1e-3
    z = torch.randn_like(x)
    n = x.numel()
    lat = 4 * n
    for name in ("ddim", "dps", "mpgd", "dsg", "diffmusic"):
        sched = dm.get_scheduler(name)(operator=None, **stubs.MUSICLDM_SCHED)
        sched.set_timesteps(500)
        c = sched._coeffs(501, 1.0)
        x0 = sched._x0(x, e, c, 0)[0]
        prev = torch.empty_like(x)
        from diffmusic_b200 import _lib
        st = _lib.stream
        if name == "ddim":
            add("sched.x0", lambda: sched._x0(x, e, c, 0), 3 * lat)
            add("sched.ddim_update", lambda: _lib.call("dm_sched_ddim_update", x.data_ptr(), x0.data_ptr(),
                                                       prev.data_ptr(), n, c["sqrt_a"], c["sqrt_b"], c["sqrt_p"],
                                                       c["sqrt_1mp"], None, st()), 3 * lat)
        elif name == "dps":
            add("sched.dps_update", lambda: _lib.call("dm_sched_dps_update", x.data_ptr(), x0.data_ptr(), g.data_ptr(),
                                                      z.data_ptr(), prev.data_ptr(), n, c["sqrt_a"], c["sqrt_b"],
                                                      c["sqrt_p"], c["dir_coef"], c["std"], 5e-4, None, st()), 5 * lat)
        elif name == "mpgd":
            x0n = torch.empty_like(x)
            add("sched.mpgd_update", lambda: _lib.call("dm_sched_mpgd_update", x.data_ptr(), x0.data_ptr(),
                                                       g.data_ptr(), z.data_ptr(), prev.data_ptr(), x0n.data_ptr(), n,
                                                       c["sqrt_a"], c["sqrt_b"], c["sqrt_p"], c["dir_coef"], c["std"],
                                                       0.005, None, st()), 6 * lat)
        elif name == "dsg":
            r = float(torch.sqrt(torch.tensor(32000.0)) * c["std"])
            add("sched.dsg_update", lambda: _lib.call("dm_sched_dsg_update", x0.data_ptr(), e.data_ptr(), g.data_ptr(),
                                                      z.data_ptr(), prev.data_ptr(), B, 32000, c["sqrt_a"], c["sqrt_p"],
                                                      c["dir_coef"], c["std"], 0.08, r, 1e-3, 1e-8, None, st()),
                5 * lat, "one 8-CTA cluster per clip, single read")
        else:
            add("sched.diffmusic_update", lambda: _lib.call("dm_sched_diffmusic_update", x0.data_ptr(), e.data_ptr(),
                                                            g.data_ptr(), z.data_ptr(), prev.data_ptr(), B, 32000,
                                                            c["sqrt_a"], c["sqrt_p"], c["dir_coef"], c["std"], 0.08,
                                                            1e-3, 1e-8, 0.9995, None, st()),
                5 * lat, "one 8-CTA cluster per clip, single read")
    # phase-preserving export (pipeline_musicldm.py:263-301): mel + shared phase in, waveform out
    mel = torch.rand(B, 1, T, M, device=dev) * 6.0 - 1.0
    ph = (torch.rand(1, F, T, device=dev) * 2.0 - 1.0) * 3.14159
    add("export.mel_to_waveform_with_phase",
        lambda: dm.mel_spectrogram_to_waveform_with_phase(mel, ph, original_waveform_length=L),
        B * (4 * M * T + 4 * L) + 4 * F * T, "InverseMelScale + istft fused; phase shared by the batch")
    if a.torch_yardstick:
        import torchaudio
        inv = torchaudio.transforms.InverseMelScale(n_stft=F, n_mels=M, sample_rate=16000).to(dev)
        w = (inv.fb @ torch.linalg.inv(inv.fb.T @ inv.fb))

        def eager():  # the same chain in eager torch on this GPU, lstsq replaced by its closed form (gels on CUDA
            lin = torch.relu(w @ mel.squeeze(1).permute(0, 2, 1))  # needs a full-rank TALL system)
            return torch.istft(lin * torch.exp(1j * ph.squeeze(0)), n_fft=1024, hop_length=160, win_length=1024)
        add("export.mel_to_waveform_with_phase[torch eager]", eager, B * (4 * M * T + 4 * L) + 4 * F * T,
            "yardstick: matmul + exp + cuFFT irfft + fold")
    print(json.dumps({"what": "operator HBM GB/s vs peak", "batch": B, "clip_samples": L, "hbm_peak_GBps": pk,
                      "l2": "flushed before every timed call", "rows": rows}, indent=1))


if __name__ == "__main__":
    main()
