"""Generate tests/golden/*.npz by running the REFERENCE's own Python (read-only, from /root/reference).

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

* operators come from /root/reference/diffmusic/inverse_problem/operator.py unmodified; their ctors hard-code
  `.to("cuda")` (operator.py:33,83,149,191,226), made a no-op here because the container has no GPU;
* schedulers come from /root/reference/diffmusic/schedulers/*.py unmodified, importing `diffusers` from the
  throw-away shim in tests/golden/_shim (base class = oracle/ddim_base.py; diffusers is not installed);
* vae / vocoder are the deterministic stubs of tests/stubs.py; inputs are the seeded synthetic tensors of
  tests/stubs.py, so the fixtures store OUTPUTS only (plus the library constants the product must reproduce).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, "_shim"))
sys.path.append("/root/reference")

from tests import stubs  # noqa: E402

_orig_to = torch.nn.Module.to


def _to_no_cuda(self, *a, **k):
    if a and isinstance(a[0], str) and a[0].startswith("cuda") and not torch.cuda.is_available():
        return self
    return _orig_to(self, *a, **k)


torch.nn.Module.to = _to_no_cuda

from diffmusic.inverse_problem import get_noiser  # noqa: E402  (reference)
from diffmusic.inverse_problem import operator as refop  # noqa: E402  (reference)
from diffmusic.schedulers import get_scheduler  # noqa: E402  (reference)

assert refop.__file__.startswith("/root/reference"), refop.__file__

torch.set_num_threads(8)
L1 = 16000  # 1 s clips -> 101 frames
LP = 4000   # phase wav-space case -> 26 frames


def npf(t):
    return t.detach().cpu().numpy()


def loss_and_grad(op, wav, meas, space):
    w = wav.clone().requires_grad_(True)
    pred = op.forward(w)
    if space == "wav_form":
        diff = meas - pred
    else:
        diff = op.transform(meas) - op.transform(pred)
    loss = torch.linalg.norm(diff)
    (g,) = torch.autograd.grad(loss, w)
    return loss.detach(), g


def make_operators():
    out = {}
    noiser = get_noiser("gaussian", 0.0)
    ident = refop.IdentityOperator(sample_rate=16000)
    out["hann_window"] = npf(ident.wav2mel[0].spectrogram.window)
    fb = ident.wav2mel[0].mel_scale.fb
    nz = fb.nonzero()
    out["fb_shape"] = np.array(fb.shape)
    out["fb_idx"] = npf(nz).astype(np.int32)
    out["fb_val"] = npf(fb[nz[:, 0], nz[:, 1]])

    # masks (bit-exact index math) ------------------------------------------------------------
    kw = dict(audio_length_in_s=10, sample_rate=16000, mask_percentage=0.3, interval_s=1, mask_duration_s=0.1,
              noiser=noiser)
    box = refop.MusicInpaintingOperator(mask_type="box", start_inpainting_s=2, end_inpainting_s=3, **kw)
    out["mask_box_10s"] = np.packbits(npf(box.mask[0]).astype(np.uint8))
    box_f = refop.MusicInpaintingOperator(mask_type="box", start_inpainting_s=1.37, end_inpainting_s=2.913, **kw)
    out["mask_boxfrac_10s"] = np.packbits(npf(box_f.mask[0]).astype(np.uint8))
    torch.manual_seed(7)
    rnd = refop.MusicInpaintingOperator(mask_type="random", start_inpainting_s=None, end_inpainting_s=None, **kw)
    out["mask_random_seed7_10s"] = np.packbits(npf(rnd.mask[0]).astype(np.uint8))
    per = refop.MusicInpaintingOperator(mask_type="periodic", start_inpainting_s=None, end_inpainting_s=None, **kw)
    out["mask_periodic_10s"] = np.packbits(npf(per.mask[0]).astype(np.uint8))

    # operator outputs on seeded 1 s clips (B = 2) ----------------------------------------------
    wav = stubs.synth_clips(2, L1)           # "prediction"
    ref = stubs.synth_clips(2, L1, first=50)  # source of the measurement
    kw1 = dict(kw, audio_length_in_s=1)
    inp = refop.MusicInpaintingOperator(mask_type="box", start_inpainting_s=0.25, end_inpainting_s=0.5, **kw1)
    out["identity_transform"] = npf(ident.transform(wav))
    out["inpaint_forward"] = npf(inp.forward(wav))
    out["inpaint_transform"] = npf(inp.transform(inp.forward(wav)))
    for space in ("mel_spectrogram", "wav_form"):
        l, g = loss_and_grad(inp, wav[:1], inp.forward(ref[:1]), space)
        out[f"inpaint_{space}_loss"], out[f"inpaint_{space}_grad"] = npf(l), npf(g)
    l, g = loss_and_grad(ident, wav[:1], ref[:1], "mel_spectrogram")
    out["identity_mel_spectrogram_loss"], out["identity_mel_spectrogram_grad"] = npf(l), npf(g)

    for scale in (2, 10):
        sr = refop.SuperResolutionOperator(sample_rate=16000, scale=scale, noiser=noiser)
        out[f"resample_kernel_s{scale}"] = npf(sr.resampler.kernel)
        out[f"resample_width_s{scale}"] = np.array(sr.resampler.width)
        out[f"superres_forward_s{scale}"] = npf(sr.forward(wav))
        out[f"superres_transform_s{scale}"] = npf(sr.transform(sr.forward(wav)))
        for space in ("mel_spectrogram", "wav_form"):
            l, g = loss_and_grad(sr, wav[:1], sr.forward(ref[:1]), space)
            out[f"superres_s{scale}_{space}_loss"], out[f"superres_s{scale}_{space}_grad"] = npf(l), npf(g)

    # dereverberation: the reference redraws the IR inside forward from the global CPU generator
    for K, decay in ((800, 0.85), (5000, 0.99), (801, 0.9)):
        dv = refop.MusicDereverberationOperator(ir_length=K, decay_factor=decay, noiser=noiser)
        torch.manual_seed(100 + K)
        ir = dv.generate_impulse_response(ir_length=K, decay_factor=decay)
        out[f"dereverb_ir_K{K}"] = npf(ir)
        torch.manual_seed(100 + K)
        out[f"dereverb_forward_K{K}"] = npf(dv.forward(wav))
        torch.manual_seed(100 + K)
        meas = dv.forward(ref[:1])
        for space in ("mel_spectrogram", "wav_form"):
            torch.manual_seed(100 + K)
            l, g = loss_and_grad(dv, wav[:1], meas, space)
            out[f"dereverb_K{K}_{space}_loss"], out[f"dereverb_K{K}_{space}_grad"] = npf(l), npf(g)

    ph = refop.PhaseRetrievalOperator(n_fft=1024, hop_length=160, win_length=1024, noiser=noiser)
    wp, rp = wav[:, :LP], ref[:, :LP]
    out["phase_forward"] = npf(ph.forward(wp))
    out["phase_transform"] = npf(ph.transform(ph.forward(wp)))
    for space in ("mel_spectrogram", "wav_form"):
        l, g = loss_and_grad(ph, wp[:1], ph.forward(rp[:1]), space)
        out[f"phase_{space}_loss"], out[f"phase_{space}_grad"] = npf(l), npf(g)

    # gaussian noiser with sigma > 0 (noise.py:13-18): global CPU generator
    nz = get_noiser("gaussian", 0.05)
    torch.manual_seed(5)
    out["gaussian_noise_s0.05_seed5"] = npf(nz(wav[:1, :256]))
    np.savez_compressed(os.path.join(HERE, "operators.npz"), **out)
    print("operators.npz:", len(out), "arrays")


def make_steps():
    """One batch-1 reference `.step` per (scheduler, operator, t, space); 1 s clips, latent (1, 8, 25, 16)."""
    out = {}
    noiser = get_noiser("gaussian", 0.0)
    vae, voc = stubs.StubVAE(), stubs.StubVocoder()
    ref_wav = stubs.synth_clips(1, L1, first=50)
    kw1 = dict(audio_length_in_s=1, sample_rate=16000, mask_percentage=0.3, interval_s=1, mask_duration_s=0.1,
               noiser=noiser)
    ops = {
        "inpainting": refop.MusicInpaintingOperator(mask_type="box", start_inpainting_s=0.25, end_inpainting_s=0.5,
                                                    **kw1),
        "super_resolution": refop.SuperResolutionOperator(sample_rate=16000, scale=2, noiser=noiser),
        "phase_retrieval": refop.PhaseRetrievalOperator(noiser=noiser),
        "dereverberation": refop.MusicDereverberationOperator(ir_length=800, decay_factor=0.85, noiser=noiser),
        "identity": refop.IdentityOperator(sample_rate=16000),
    }
    cases = [  # scheduler, operator, eta, rate, space
        ("ddim", "inpainting", 0.0, None, "mel_spectrogram"),
        ("dps", "super_resolution", 0.0, 5e-4, "mel_spectrogram"),
        ("dps", "inpainting", 0.5, 5e-4, "wav_form"),
        ("mpgd", "identity", 0.0, 0.005, "mel_spectrogram"),
        ("mpgd", "inpainting", 1.0, 0.005, "mel_spectrogram"),
        ("dsg", "phase_retrieval", 1.0, 0.08, "mel_spectrogram"),
        ("dsg", "phase_retrieval", 1.0, 0.08, "wav_form"),
        ("dsg", "inpainting", 1.0, 0.08, "mel_spectrogram"),
        ("diffmusic", "dereverberation", 1.0, 0.08, "mel_spectrogram"),
        ("diffmusic", "inpainting", 1.0, 0.08, "mel_spectrogram"),
        ("diffmusic", "super_resolution", 1.0, 0.08, "wav_form"),
    ]
    x, e = stubs.synth_latents(1, 25)
    for sched_name, op_name, eta, rate, space in cases:
        op = ops[op_name]
        sched = get_scheduler(sched_name)(operator=op, **stubs.MUSICLDM_SCHED)
        sched.set_timesteps(500)
        torch.manual_seed(321)  # dereverb IR for the measurement
        meas = op.forward(ref_wav)
        for t in (999, 501, 1):
            gen = torch.Generator().manual_seed(3000)
            kwargs = dict(eta=eta, generator=gen, measurement=meas, vae=vae, vocoder=voc,
                          original_waveform_length=L1, supervised_space=space)
            if rate is not None:
                kwargs["ip_guidance_rate"] = rate
            if sched_name == "ddim":  # reference DDIM needs these to be tensors (scheduling_ddim.py:102-103)
                kwargs.update(encoder_hidden_states=torch.zeros(1), encoder_hidden_states_1=torch.zeros(1))
                kwargs.pop("supervised_space")
            torch.manual_seed(654 + t)  # dereverb IR drawn inside the step
            o = sched.step(e, t, x, **kwargs)
            key = f"{sched_name}|{op_name}|{space}|eta{eta}|t{t}"
            out[key + "|prev"] = npf(o.prev_sample)
            out[key + "|x0"] = npf(o.pred_original_sample)
            out[key + "|loss"] = npf(o.loss.float())
    # scheduler constants
    s = get_scheduler("dps")(operator=None, **stubs.MUSICLDM_SCHED)
    s.set_timesteps(500)
    out["timesteps_500"] = npf(s.timesteps)
    out["alphas_cumprod"] = npf(s.alphas_cumprod)
    out["final_alpha_cumprod"] = npf(s.final_alpha_cumprod)
    np.savez_compressed(os.path.join(HERE, "steps.npz"), **out)
    print("steps.npz:", len(out), "arrays")


def make_steps_clip():
    """clip_sample=True (the constructor default, scheduling_dps.py:29): the reference differentiates through the base
    step's `pred_original_sample.clamp(+-clip_sample_range)`, so autograd zeroes the guidance gradient wherever the
    unclipped x0 lies outside the range (DPS / DSG / DiffMusic re-leaf `sample`; MPGD leafs the clipped x0).
    Seeded latents give |x0| > 1 for almost every element at t = 999, about half at t = 501, a third at t = 1."""
    out = {}
    noiser = get_noiser("gaussian", 0.0)
    vae, voc = stubs.StubVAE(), stubs.StubVocoder()
    ref_wav = stubs.synth_clips(1, L1, first=50)
    op = refop.MusicInpaintingOperator(audio_length_in_s=1, sample_rate=16000, mask_percentage=0.3, interval_s=1,
                                       mask_duration_s=0.1, noiser=noiser, mask_type="box", start_inpainting_s=0.25,
                                       end_inpainting_s=0.5)
    cfg = dict(stubs.MUSICLDM_SCHED, clip_sample=True)
    x, e = stubs.synth_latents(1, 25)
    meas = op.forward(ref_wav)
    for sched_name, eta, rate in (("dps", 0.0, 5e-4), ("mpgd", 0.0, 0.005), ("dsg", 1.0, 0.08),
                                  ("diffmusic", 1.0, 0.08)):
        sched = get_scheduler(sched_name)(operator=op, **cfg)
        sched.set_timesteps(500)
        for t in (999, 501, 1):
            gen = torch.Generator().manual_seed(3000)
            o = sched.step(e, t, x, eta=eta, generator=gen, measurement=meas, vae=vae, vocoder=voc,
                           original_waveform_length=L1, supervised_space="mel_spectrogram", ip_guidance_rate=rate)
            key = f"{sched_name}|inpainting|mel_spectrogram|eta{eta}|t{t}"
            out[key + "|prev"] = npf(o.prev_sample)
            out[key + "|x0"] = npf(o.pred_original_sample)
            out[key + "|loss"] = npf(o.loss.float())
            out[key + "|clipped_frac"] = np.float32((npf(o.pred_original_sample).__abs__() >= 1.0).mean())
    np.savez_compressed(os.path.join(HERE, "steps_clip.npz"), **out)
    print("steps_clip.npz:", len(out), "arrays")


if __name__ == "__main__":
    if "--clip-only" in sys.argv:
        make_steps_clip()
        sys.exit(0)
    make_operators()
    make_steps()
    make_steps_clip()
