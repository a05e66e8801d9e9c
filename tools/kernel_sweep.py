"""A/B of kernel-selection knobs (dm_set_tuning): device time per C-ABI call and bit-equality of the results.

    python tools/kernel_sweep.py [--batch 16] > gpurun_out/kernel_sweep_b16.json

Every chain runs behind an L2 flush with a CUDA-event pair around each C-ABI call (the flush keeps the host ahead of the
device, so a bracket holds the kernel, not its launch latency); median of `iters` per call name.
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
import diffmusic_b200.operators as ops_mod  # noqa: E402
from diffmusic_b200 import _lib  # noqa: E402
from tests import stubs  # noqa: E402

L = 160000
TUNE_STREAM, TUNE_PDL = 0, 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    B, dev = a.batch, torch.device("cuda", 0)
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    events = {}
    orig = _lib.call
    state = {"on": False}

    def timed_call(name, *args):
        if not state["on"]:
            return orig(name, *args)
        s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        r = orig(name, *args)
        t.record()
        events.setdefault(name, []).append((s, t))
        return r

    ops_mod._lib.call = timed_call

    def run(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        events.clear()
        whole = []
        for i in range(a.iters):
            for j in range(4):
                flush.fill_(float(i + j))
            s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            state["on"] = True
            s.record()
            out = fn()
            t.record()
            state["on"] = False
            whole.append((s, t))
        torch.cuda.synchronize()
        row = {k: round(1e3 * statistics.median(s.elapsed_time(t) for s, t in v), 2) for k, v in events.items()}
        row["chain_us"] = round(1e3 * statistics.median(s.elapsed_time(t) for s, t in whole), 2)
        return row, out

    wav = (0.1 * torch.randn(B, L, device=dev)).contiguous()
    ref_wav = stubs.synth_clips(1, L, first=50).to(dev)
    nz = dm.get_noiser("gaussian", 0.0)
    sr = dm.SuperResolutionOperator(16000, scale=2, noiser=nz)
    inp = dm.MusicInpaintingOperator(10, 16000, "box", 2, 3, 0.3, 0.1, 1, noiser=nz)
    chains = {
        "super_resolution.forward": lambda: sr.forward(wav),
        "super_resolution.loss+vjp[mel]": lambda m=sr.forward(ref_wav): sr.fused_loss_and_grad(wav, m, "mel_spectrogram"),
        "super_resolution.loss+vjp[wav]": lambda m=sr.forward(ref_wav): sr.fused_loss_and_grad(wav, m, "wav_form"),
        "inpainting.loss+vjp[mel]": lambda m=inp.forward(ref_wav): inp.fused_loss_and_grad(wav, m, "mel_spectrogram"),
        "inpainting.loss+vjp[wav]": lambda m=inp.forward(ref_wav): inp.fused_loss_and_grad(wav, m, "wav_form"),
    }
    out = {"batch": B, "stream_kernels": {}}
    for name, fn in chains.items():
        res = {}
        for knob, key in ((0, 0), (2, 1)):  # plain kernels / persistent kernels whatever the batch
            _lib.call("dm_set_tuning", TUNE_STREAM, knob)
            row, val = run(fn)
            res[str(key)] = row
            res[f"val{key}"] = val
        _lib.call("dm_set_tuning", TUNE_STREAM, 1)
        v0, v1 = res.pop("val0"), res.pop("val1")
        v0, v1 = (v0 if isinstance(v0, tuple) else (v0,)), (v1 if isinstance(v1, tuple) else (v1,))
        res["bit_identical"] = all(torch.equal(p, q) for p, q in zip(v0, v1))
        out["stream_kernels"][name] = res

    # programmatic dependent launch along the chains (knob 1)
    der = dm.MusicDereverberationOperator(ir_length=5000, decay_factor=0.99, noiser=nz)
    torch.manual_seed(0)
    chains["dereverberation.loss+vjp[mel]"] = lambda m=der.forward(ref_wav): der.fused_loss_and_grad(wav, m, "mel_spectrogram")
    out["pdl"] = {}
    for name in ("super_resolution.loss+vjp[mel]", "inpainting.loss+vjp[mel]", "dereverberation.loss+vjp[mel]"):
        res = {}
        for knob in (0, 1):
            _lib.call("dm_set_tuning", TUNE_PDL, knob)
            torch.manual_seed(1)
            row, val = run(chains[name])
            res[str(knob)] = row
            res[f"val{knob}"] = val
        _lib.call("dm_set_tuning", TUNE_PDL, 0)
        v0, v1 = res.pop("val0"), res.pop("val1")
        res["bit_identical"] = all(torch.equal(p, q) for p, q in zip(v0, v1))
        out["pdl"][name] = res
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
