"""GPU parity tests (run on the B200 box with `-m gpu`): the CUDA path, called through the C ABI by the host-side mirror,
against (a) the committed golden fixtures produced by the reference's own Python and (b) the CPU oracle on the same
seeded inputs.  Tolerance: <= 1e-4 relative L2 for operator outputs, gradients and per-step latents in fp32
(BASELINE.json north_star); masks / indices bit-exact (tests/test_abi_and_host.py)."""
import numpy as np
import pytest
import torch

import diffmusic_b200 as dm
from oracle import operators as oo
from oracle import steps as osteps
from tests import stubs
from tests.conftest import rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-4
L1, LP = 16000, 4000
DEV = "cuda"


def _noiser():
    return dm.get_noiser("gaussian", 0.0)


def _inpaint(seconds=1, start=0.25, end=0.5):
    return dm.MusicInpaintingOperator(audio_length_in_s=seconds, sample_rate=16000, mask_type="box",
                                      start_inpainting_s=start, end_inpainting_s=end, mask_percentage=0.3,
                                      mask_duration_s=0.1, interval_s=1, noiser=_noiser())


def _loss_grad(op, wav, meas, space):
    w = wav.clone().requires_grad_(True)
    loss = op.guidance_loss(w, meas, space)
    (g,) = torch.autograd.grad(loss.sum(), w)
    return loss.detach(), g


# ------------------------------------------------------------------------------------------------ operators vs golden
def test_forward_and_transform_vs_golden(golden_ops):
    wav = stubs.synth_clips(2, L1).to(DEV)
    ident, inp = dm.IdentityOperator(16000), _inpaint()
    assert rel_l2(ident.transform(wav), golden_ops["identity_transform"]) < TOL
    assert ident.forward(wav) is wav
    y = inp.forward(wav)
    assert np.array_equal(y.cpu().numpy(), golden_ops["inpaint_forward"])  # x * {0,1}: exact
    assert rel_l2(inp.transform(y), golden_ops["inpaint_transform"]) < TOL
    for s in (2, 10):
        sr = dm.SuperResolutionOperator(sample_rate=16000, scale=s, noiser=_noiser())
        y = sr.forward(wav)
        assert y.shape == golden_ops[f"superres_forward_s{s}"].shape
        assert rel_l2(y, golden_ops[f"superres_forward_s{s}"]) < TOL
        assert rel_l2(sr.transform(y), golden_ops[f"superres_transform_s{s}"]) < TOL
    ph = dm.PhaseRetrievalOperator(noiser=_noiser())
    mag = ph.forward(wav[:, :LP])
    assert mag.shape == golden_ops["phase_forward"].shape
    assert rel_l2(mag, golden_ops["phase_forward"]) < TOL
    assert rel_l2(ph.transform(mag), golden_ops["phase_transform"]) < TOL
    for K, decay in ((800, 0.85), (5000, 0.99), (801, 0.9)):
        dv = dm.MusicDereverberationOperator(ir_length=K, decay_factor=decay, noiser=_noiser())
        torch.manual_seed(100 + K)
        y = dv.forward(wav)
        assert np.array_equal(dv.last_ir.numpy(), golden_ops[f"dereverb_ir_K{K}"])
        assert y.shape == golden_ops[f"dereverb_forward_K{K}"].shape
        assert rel_l2(y, golden_ops[f"dereverb_forward_K{K}"]) < TOL


def test_forward_accepts_cpu_tensors_like_run_py(golden_ops):
    """run.py:286,312 builds the measurement from CPU tensors; result comes back on the CPU."""
    wav = stubs.synth_clips(2, L1)
    sr = dm.SuperResolutionOperator(sample_rate=16000, scale=2, noiser=_noiser())
    y = sr.forward(wav)
    assert y.device.type == "cpu" and rel_l2(y, golden_ops["superres_forward_s2"]) < TOL
    m = dm.IdentityOperator(16000).transform(wav)
    assert m.device.type == "cpu" and rel_l2(m, golden_ops["identity_transform"]) < TOL


def test_gaussian_noise_matches_reference_draw(golden_ops):
    wav = stubs.synth_clips(1, L1)[:, :256]
    torch.manual_seed(5)
    out = dm.get_noiser("gaussian", 0.05)(wav)
    assert np.array_equal(out.numpy(), golden_ops["gaussian_noise_s0.05_seed5"])
    # CUDA tensors: same torch draw, added by dm_add_scaled
    torch.manual_seed(5)
    n = torch.randn_like(wav.to(DEV))
    torch.manual_seed(5)
    out = dm.get_noiser("gaussian", 0.05)(wav.to(DEV))
    assert rel_l2(out, wav.to(DEV) + n * 0.05) < 1e-7


@pytest.mark.parametrize("space", ["mel_spectrogram", "wav_form"])
def test_loss_and_vjp_vs_golden(golden_ops, space):
    wav = stubs.synth_clips(1, L1).to(DEV)
    ref = stubs.synth_clips(1, L1, first=50).to(DEV)
    ops = {"inpaint": _inpaint(),
           "superres_s2": dm.SuperResolutionOperator(16000, scale=2, noiser=_noiser()),
           "superres_s10": dm.SuperResolutionOperator(16000, scale=10, noiser=_noiser())}
    if space == "mel_spectrogram":
        ops["identity"] = dm.IdentityOperator(16000)
    for name, op in ops.items():
        loss, g = _loss_grad(op, wav, op.forward(ref), space)
        gl = float(golden_ops[f"{name}_{space}_loss"])
        assert abs(float(loss) - gl) <= TOL * abs(gl), name
        assert rel_l2(g, golden_ops[f"{name}_{space}_grad"]) < TOL, name
    for K, decay in ((800, 0.85), (5000, 0.99), (801, 0.9)):
        dv = dm.MusicDereverberationOperator(ir_length=K, decay_factor=decay, noiser=_noiser())
        torch.manual_seed(100 + K)
        meas = dv.forward(ref)
        torch.manual_seed(100 + K)
        loss, g = _loss_grad(dv, wav, meas, space)
        gl = float(golden_ops[f"dereverb_K{K}_{space}_loss"])
        assert abs(float(loss) - gl) <= TOL * abs(gl), K
        assert rel_l2(g, golden_ops[f"dereverb_K{K}_{space}_grad"]) < TOL, K
    ph = dm.PhaseRetrievalOperator(noiser=_noiser())
    loss, g = _loss_grad(ph, wav[:, :LP], ph.forward(ref[:, :LP]), space)
    gl = float(golden_ops[f"phase_{space}_loss"])
    assert abs(float(loss) - gl) <= TOL * abs(gl)
    assert rel_l2(g, golden_ops[f"phase_{space}_grad"]) < TOL


def test_loss_only_path_matches():
    """no_grad evaluation (loss only) equals the loss of the fused forward+VJP launch."""
    wav = stubs.synth_clips(2, L1).to(DEV)
    ref = stubs.synth_clips(1, L1, first=50).to(DEV)
    for op in (_inpaint(), dm.SuperResolutionOperator(16000, 2, _noiser()),
               dm.MusicDereverberationOperator(800, 0.85, _noiser())):
        torch.manual_seed(1)
        meas = op.forward(ref)
        torch.manual_seed(2)
        with torch.no_grad():
            l0 = op.guidance_loss(wav, meas, "mel_spectrogram")
        torch.manual_seed(2)
        l1, _ = _loss_grad(op, wav, meas, "mel_spectrogram")
        assert torch.equal(l0, l1)


# ------------------------------------------------------------------------------------------------ scheduler steps
def _ops():
    return {"inpainting": _inpaint(),
            "super_resolution": dm.SuperResolutionOperator(sample_rate=16000, scale=2, noiser=_noiser()),
            "phase_retrieval": dm.PhaseRetrievalOperator(noiser=_noiser()),
            "dereverberation": dm.MusicDereverberationOperator(ir_length=800, decay_factor=0.85, noiser=_noiser()),
            "identity": dm.IdentityOperator(sample_rate=16000)}


RATES = {"ddim": None, "dps": 5e-4, "mpgd": 0.005, "dsg": 0.08, "diffmusic": 0.08}


def test_steps_vs_reference_golden(golden_steps):
    """all 33 (scheduler, operator, space, eta, t) cases the reference's own scheduler files were run on."""
    vae, voc = stubs.StubVAE().to(DEV), stubs.StubVocoder().to(DEV)
    ref_wav = stubs.synth_clips(1, L1, first=50).to(DEV)
    x, e = stubs.synth_latents(1, 25)
    x, e = x.to(DEV), e.to(DEV)
    ops = _ops()
    cases = [k[:-5] for k in golden_steps.files if k.endswith("|prev")]
    assert len(cases) == 33
    worst = 0.0
    for key in cases:
        sched_name, op_name, space, eta, t = key.split("|")
        eta, t = float(eta[3:]), int(t[1:])
        op = ops[op_name]
        sched = dm.get_scheduler(sched_name)(operator=op, **stubs.MUSICLDM_SCHED)
        sched.set_timesteps(500)
        torch.manual_seed(321)
        meas = op.forward(ref_wav)
        kwargs = dict(eta=eta, generator=torch.Generator().manual_seed(3000), measurement=meas, vae=vae, vocoder=voc,
                      original_waveform_length=L1)
        if RATES[sched_name] is not None:
            kwargs.update(ip_guidance_rate=RATES[sched_name], supervised_space=space)
        torch.manual_seed(654 + t)
        out = sched.step(e, t, x, **kwargs)
        ep, e0 = rel_l2(out.prev_sample, golden_steps[key + "|prev"]), rel_l2(out.pred_original_sample,
                                                                               golden_steps[key + "|x0"])
        worst = max(worst, ep, e0)
        assert ep < TOL and e0 < TOL, (key, ep, e0)
        gl = float(golden_steps[key + "|loss"].ravel()[0])
        assert abs(float(out.loss.float().ravel()[0]) - gl) <= TOL * max(1.0, abs(gl)), key
        assert not out.prev_sample.requires_grad and out.loss.numel() == 1
    print("worst rel-L2 over 33 reference step cases:", worst)


def test_clip_sample_steps_vs_reference_golden(golden_steps_clip):
    """clip_sample=True (constructor default): the guidance gradient is masked where the unclipped x0 leaves
    +-clip_sample_range, as autograd does through the reference's clamp (scheduling_dps.py:165-175); 12 cases produced
    by the reference's scheduler files, eager and CUDA-graph replay."""
    from diffmusic_b200.graph import GraphedGuidedStep
    vae, voc = stubs.StubVAE().to(DEV), stubs.StubVocoder().to(DEV)
    ref_wav = stubs.synth_clips(1, L1, first=50).to(DEV)
    x, e = stubs.synth_latents(1, 25)
    x, e = x.to(DEV), e.to(DEV)
    op = _inpaint()
    meas = op.forward(ref_wav)
    cases = [k[:-5] for k in golden_steps_clip.files if k.endswith("|prev")]
    assert len(cases) == 12
    for key in cases:
        sched_name, _, space, eta, t = key.split("|")
        eta, t = float(eta[3:]), int(t[1:])
        sched = dm.get_scheduler(sched_name)(operator=op, **dict(stubs.MUSICLDM_SCHED, clip_sample=True))
        sched.set_timesteps(500)
        kwargs = dict(eta=eta, measurement=meas, vae=vae, vocoder=voc, original_waveform_length=L1,
                      ip_guidance_rate=RATES[sched_name], supervised_space=space)
        out = sched.step(e, t, x, generator=torch.Generator().manual_seed(3000), **kwargs)
        ep = rel_l2(out.prev_sample, golden_steps_clip[key + "|prev"])
        e0 = rel_l2(out.pred_original_sample, golden_steps_clip[key + "|x0"])
        assert ep < TOL and e0 < TOL, (key, ep, e0)
        gl = float(golden_steps_clip[key + "|loss"].ravel()[0])
        assert abs(float(out.loss.float().ravel()[0]) - gl) <= TOL * max(1.0, abs(gl)), key
    # graph replay keeps the mask per timestep (the scalars come from the device-side coefficient row)
    sched = dm.DPSScheduler(operator=op, **dict(stubs.MUSICLDM_SCHED, clip_sample=True))
    sched.set_timesteps(500)
    graphed = GraphedGuidedStep(sched, x.shape, measurement=meas, vae=vae, vocoder=voc, original_waveform_length=L1,
                                ip_guidance_rate=RATES["dps"], eta=0.0, supervised_space="mel_spectrogram")
    for t in (999, 501, 1):
        out = graphed(e, t, x)
        key = f"dps|inpainting|mel_spectrogram|eta0.0|t{t}"
        assert rel_l2(out.prev_sample, golden_steps_clip[key + "|prev"]) < TOL, key


@pytest.mark.parametrize("sched_name,op_name,eta", [("dps", "super_resolution", 0.0), ("mpgd", "inpainting", 1.0),
                                                    ("dsg", "phase_retrieval", 1.0), ("diffmusic", "inpainting", 1.0),
                                                    ("diffmusic", "dereverberation", 1.0)])
def test_batched_step_is_per_clip(sched_name, op_name, eta):
    """B = 3 with a list of generators == three independent batch-1 oracle steps."""
    B = 3
    vae, voc = stubs.StubVAE(), stubs.StubVocoder()
    ref_wav = stubs.synth_clips(B, L1, first=50)
    x, e = stubs.synth_latents(B, 25)
    mask = oo.inpaint_mask(1, 16000, "box", 0.25, 0.5)
    ir = None
    if op_name == "dereverberation":
        torch.manual_seed(9)
        ir = oo.draw_impulse_response(800, 0.85)
    oops = {"inpainting": oo.OracleOperator("inpainting", mask=mask),
            "super_resolution": oo.OracleOperator("super_resolution", scale=2),
            "phase_retrieval": oo.OracleOperator("phase_retrieval"),
            "dereverberation": oo.OracleOperator("dereverberation", fixed_ir=ir)}
    oop = oops[op_name]
    meas = oop.forward(ref_wav)  # per-clip measurements (B, ...)
    base = osteps.make_base(**stubs.MUSICLDM_SCHED)
    base.set_timesteps(500)
    t = 501
    want = osteps.per_clip_step(sched_name, base, oop, e, t, x, generators=stubs.step_generators(B),
                                measurement=meas, eta=eta, ip_guidance_rate=RATES[sched_name], vae=vae, vocoder=voc,
                                original_waveform_length=L1, supervised_space="mel_spectrogram")
    op = _ops()[op_name]
    if ir is not None:
        op.generate_impulse_response = lambda ir_length, decay_factor: ir  # same IR as the oracle
    sched = dm.get_scheduler(sched_name)(operator=op, **stubs.MUSICLDM_SCHED)
    sched.set_timesteps(500)
    got = sched.step(e.to(DEV), t, x.to(DEV), eta=eta, generator=stubs.step_generators(B),
                     measurement=meas.to(DEV), vae=vae.to(DEV), vocoder=voc.to(DEV), original_waveform_length=L1,
                     ip_guidance_rate=RATES[sched_name], supervised_space="mel_spectrogram")
    assert rel_l2(got.prev_sample, want.prev_sample) < TOL
    assert rel_l2(got.pred_original_sample, want.pred_original_sample) < TOL
    assert rel_l2(got.loss_per_clip, want.loss) < TOL
    assert abs(float(got.loss) - float(torch.linalg.norm(want.loss))) < TOL * float(torch.linalg.norm(want.loss))
    for i in range(B):  # every clip individually, not just in aggregate
        assert rel_l2(got.prev_sample[i], want.prev_sample[i]) < TOL


@pytest.mark.parametrize("cfg", ["cfg1", "cfg2", "cfg3", "cfg4"])
def test_step_parity_at_baseline_shapes(cfg):
    """`.step` at the exact BASELINE.json shapes -- 10 s clips (L = 160000), latents (B, 8, 250, 16), the bench's batch
    sizes -- against `oracle.steps.per_clip_step` (scheduling_dps.py:137-219 and its clones), t in {999, 501, 1}:
      cfg1 inpainting box[2 s, 3 s) + DDIM, B = 1;      cfg2 super-resolution x2 + DPS, B = 16;
      cfg3 phase retrieval + DSG (eta 1), B = 8;        cfg4 dereverberation K = 5000 + DiffMusic (eta 1), B = 16.
    The product runs the whole batch in one call; the oracle runs the first, a middle and the last clip as batch-1
    steps (per-clip semantics; the CPU convolution with 5000 taps costs seconds per clip)."""
    sched_name, op_name, B, eta = {"cfg1": ("ddim", "inpainting", 1, 0.0), "cfg2": ("dps", "super_resolution", 16, 0.0),
                                   "cfg3": ("dsg", "phase_retrieval", 8, 1.0),
                                   "cfg4": ("diffmusic", "dereverberation", 16, 1.0)}[cfg]
    L, H = 160000, 250
    vae, voc = stubs.StubVAE(), stubs.StubVocoder()
    x, e = stubs.synth_latents(B, H)
    ref = stubs.synth_clips(1, L, first=50)
    ir = None
    if op_name == "inpainting":
        oop, op = oo.OracleOperator("inpainting", mask=oo.inpaint_mask(10, 16000, "box", 2, 3)), _inpaint(10, 2, 3)
    elif op_name == "super_resolution":
        oop, op = oo.OracleOperator("super_resolution", scale=2), dm.SuperResolutionOperator(16000, 2, _noiser())
    elif op_name == "phase_retrieval":
        oop, op = oo.OracleOperator("phase_retrieval"), dm.PhaseRetrievalOperator(noiser=_noiser())
    else:
        torch.manual_seed(5)
        ir = oo.draw_impulse_response(5000, 0.99)
        oop, op = oo.OracleOperator("dereverberation", fixed_ir=ir), dm.MusicDereverberationOperator(5000, 0.99, _noiser())
        op.generate_impulse_response = lambda ir_length, decay_factor: ir  # the step's draw, same as the oracle's
    meas = oop.forward(ref)
    base = osteps.make_base(**stubs.MUSICLDM_SCHED)
    base.set_timesteps(500)
    sched = dm.get_scheduler(sched_name)(operator=op, **stubs.MUSICLDM_SCHED)
    sched.set_timesteps(500)
    vae_d, voc_d = stubs.StubVAE().to(DEV), stubs.StubVocoder().to(DEV)
    pick = sorted({0, B // 2, B - 1})
    kw = dict(eta=eta, vae=None, vocoder=None, original_waveform_length=L)
    if RATES[sched_name] is not None:
        kw.update(ip_guidance_rate=RATES[sched_name], supervised_space="mel_spectrogram")
    for t in (999, 501, 1):
        got = sched.step(e.to(DEV), t, x.to(DEV), generator=stubs.step_generators(B), measurement=meas.to(DEV),
                         **dict(kw, vae=vae_d, vocoder=voc_d))
        gens = stubs.step_generators(B)
        for i in pick:
            want = osteps.reference_step(sched_name, base, oop, e[i:i + 1], t, x[i:i + 1], generator=gens[i],
                                         measurement=meas, **dict(kw, vae=vae, vocoder=voc))
            assert rel_l2(got.prev_sample[i:i + 1], want.prev_sample) < TOL, (cfg, t, i)
            assert rel_l2(got.pred_original_sample[i:i + 1], want.pred_original_sample) < TOL, (cfg, t, i)
            if sched_name != "ddim":
                wl = float(want.loss)
                assert abs(float(got.loss_per_clip[i]) - wl) < TOL * wl, (cfg, t, i)


def test_step_argument_errors():
    op = _inpaint()
    sched = dm.DPSScheduler(operator=op, **stubs.MUSICLDM_SCHED)
    sched.set_timesteps(500)
    x, e = stubs.synth_latents(1, 25)
    vae, voc = stubs.StubVAE().to(DEV), stubs.StubVocoder().to(DEV)
    meas = op.forward(stubs.synth_clips(1, L1).to(DEV))
    kw = dict(measurement=meas, vae=vae, vocoder=voc, original_waveform_length=L1)
    with pytest.raises(ValueError):
        sched.step(e.to(DEV), 501, x.to(DEV), supervised_space="nope", **kw)
    with pytest.raises(ValueError):
        sched.step(e.to(DEV), 501, x.to(DEV), eta=1.0, generator=torch.Generator().manual_seed(0),
                   variance_noise=torch.zeros_like(x).to(DEV), **kw)
    # unknown kwargs the pipelines pass are swallowed
    out = sched.step(e.to(DEV), 501, x.to(DEV), ditto_optimizer=None, init_latents=None, **kw)
    assert torch.isfinite(out.loss)
    # DDIM accepts missing encoder_hidden_states (superset of the reference)
    d = dm.DDIMScheduler(operator=op, **stubs.MUSICLDM_SCHED)
    d.set_timesteps(500)
    o = d.step(e.to(DEV), 501, x.to(DEV), **kw)
    assert o.loss.tolist() == [501] and o.encoder_hidden_states is None


def test_nan_propagates_to_loss_like_the_reference():
    """pipeline_musicldm.py:742: the caller restarts on torch.isnan(out.loss); a NaN waveform must surface there."""
    op = _inpaint()
    wav = stubs.synth_clips(1, L1).to(DEV)
    meas = op.forward(wav)
    bad = wav.clone()
    bad[0, 9000] = float("nan")
    loss, g = _loss_grad(op, bad, meas, "mel_spectrogram")
    assert torch.isnan(loss).all()
    # zero residual -> zero gradient (torch.linalg.norm backward is masked at 0)
    loss, g = _loss_grad(op, wav, meas, "wav_form")
    assert float(loss) == 0.0 and float(g.abs().max()) == 0.0


def test_bit_reproducible():
    op = dm.SuperResolutionOperator(16000, 2, _noiser())
    wav = stubs.synth_clips(4, L1).to(DEV)
    meas = op.forward(stubs.synth_clips(1, L1, first=50).to(DEV))
    a = _loss_grad(op, wav, meas, "mel_spectrogram")
    b = _loss_grad(op, wav, meas, "mel_spectrogram")
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])


# ------------------------------------------------------------------------------------------------ full-size properties
@pytest.mark.parametrize("op_name", ["inpainting", "super_resolution", "dereverberation", "phase_retrieval"])
def test_full_size_10s_against_oracle(op_name):
    """BASELINE config shapes (10 s, L = 160000, T = 1001) against the CPU oracle: loss and gradient."""
    L, B = 160000, 2
    wav = stubs.synth_clips(B, L)
    ref = stubs.synth_clips(1, L, first=50)
    mask = oo.inpaint_mask(10, 16000, "box", 2, 3)
    torch.manual_seed(77)
    ir = oo.draw_impulse_response(5000, 0.99)
    oop = {"inpainting": oo.OracleOperator("inpainting", mask=mask),
           "super_resolution": oo.OracleOperator("super_resolution", scale=2),
           "dereverberation": oo.OracleOperator("dereverberation", fixed_ir=ir),
           "phase_retrieval": oo.OracleOperator("phase_retrieval")}[op_name]
    op = {"inpainting": _inpaint(10, 2, 3),
          "super_resolution": dm.SuperResolutionOperator(16000, 2, _noiser()),
          "dereverberation": dm.MusicDereverberationOperator(5000, 0.99, _noiser()),
          "phase_retrieval": dm.PhaseRetrievalOperator(noiser=_noiser())}[op_name]
    if op_name == "dereverberation":
        op.generate_impulse_response = lambda ir_length, decay_factor: ir
    meas = oop.forward(ref)
    assert rel_l2(op.forward(ref.to(DEV)), meas) < TOL
    for i in range(B):
        w = wav[i:i + 1].clone().requires_grad_(True)
        want = torch.linalg.norm(oop.transform(meas) - oop.transform(oop.forward(w)))
        (gw,) = torch.autograd.grad(want, w)
        loss, g = _loss_grad(op, wav[i:i + 1].to(DEV), meas.to(DEV), "mel_spectrogram")
        assert abs(float(loss) - float(want)) < TOL * float(want)
        assert rel_l2(g, gw) < TOL


def test_operator_linearity_and_adjoint_identity():
    """size-independent properties at full size: A is linear; <A x, y> == <x, A^T y> for the VJP kernels."""
    L, B = 160000, 2
    g = torch.Generator().manual_seed(4)
    x1, x2 = torch.randn(B, L, generator=g).to(DEV), torch.randn(B, L, generator=g).to(DEV)
    torch.manual_seed(3)
    ir = oo.draw_impulse_response(5000, 0.99)
    dv = dm.MusicDereverberationOperator(5000, 0.99, _noiser())
    dv.generate_impulse_response = lambda ir_length, decay_factor: ir
    for op in (_inpaint(10, 2, 3), dm.SuperResolutionOperator(16000, 2, _noiser()), dv):
        a = op.forward(x1 + 2.0 * x2)
        b = op.forward(x1) + 2.0 * op.forward(x2)
        assert rel_l2(a, b) < 1e-5
        # adjoint identity through the wav-space fused path: loss = ||m - A x||, grad = -A^T (m - A x)/loss
        m = torch.randn_like(op.forward(x1))
        loss, grad = _loss_grad(op, x1, m, "wav_form")
        d = m - op.forward(x1)
        for i in range(B):
            lhs = float((op.forward(x2)[i].double() * d[i].double()).sum())        # <A x2, d>
            rhs = float(-(x2[i].double() * grad[i].double()).sum() * float(loss[i]))  # <x2, A^T d>
            assert abs(lhs - rhs) <= 2e-4 * max(abs(lhs), abs(rhs), 1.0)


# ------------------------------------------------------------------------------------------------ FAD statistics
@pytest.mark.parametrize("engine", ["simt", "tcgen05", "tcgen05_pair"])
@pytest.mark.parametrize("n,d", [(40, 128), (1000, 512), (4990, 768), (333, 1024), (70000, 768), (2500, 200),
                                 (3000, 640), (513, 384)])
def test_fad_moments_engines(n, d, engine):
    """sum x x^T through the SIMT tile kernel, the tcgen05 + TMA kernel (one SM per 128 x 128 tile) and its
    cta_group::2 variant (a CTA pair per 256 x 128 super-tile; odd tile counts leave a phantom block) against float64
    NumPy."""
    from diffmusic_b200 import fad
    rng = np.random.default_rng(d + n)
    X = (rng.standard_normal((n, d)) * 0.7 + rng.standard_normal(d) * 0.3).astype(np.float16)
    mom = fad.EmbeddingMoments(d, engine=engine).update(torch.from_numpy(X))
    acc = mom.acc.cpu().numpy()
    X64 = X.astype(np.float64)
    assert acc[0] == n
    # column sums: float64 throughout on the SIMT / cta_group::2 paths; on the tcgen05 path they come out of the tensor
    # cores too (A^T times a block of ones, fp32 accumulation over 512-row slabs, float64 across slabs)
    assert rel_l2(acc[1:1 + d], X64.sum(0)) < (1e-6 if engine == "tcgen05" and d % 8 == 0 and d <= 768 else 1e-7)
    got, want = acc[1 + d:].reshape(d, d), X64.T @ X64
    assert rel_l2(got, want) < 2e-6, (engine, rel_l2(got, want))
    assert np.array_equal(got, got.T)


@pytest.mark.parametrize("n,d", [(40, 128), (1000, 512), (4990, 768), (333, 1024)])
def test_fad_statistics_vs_numpy(n, d):
    """calc_embd_statistics / the online merge on the SAME fp16 array the reference sees (fadtk caches embeddings as
    fp16, model_loader.py:46-48).  np.mean of an fp16 array is an fp16 array: the product returns its
    exact float64 mean rounded to fp16 once, which may differ from numpy's float32-accumulated rounding by one fp16 ulp
    in rare elements; np.cov is float64.  The online merge of the reference feeds the fp16-rounded per-file means into
    the Chan update (fadtk/utils.py:13-16, 36-40); the product sums exact raw moments instead (superseded quirk):
    bounded by the fp16 rounding of the means, 2^-11 relative per element."""
    from diffmusic_b200 import fad
    from oracle import fad as ofad
    rng = np.random.default_rng(d)
    X = (rng.standard_normal((n, d)) * 0.7 + rng.standard_normal(d) * 0.3).astype(np.float16)
    mu, cov = fad.calc_embd_statistics(X)
    wmu, wcov = ofad.calc_embd_statistics(X)          # same fp16 input: fp16 mean, float64 covariance
    assert mu.dtype == wmu.dtype == np.float16
    ulp = np.abs(mu.view(np.int16).astype(np.int32) - wmu.view(np.int16).astype(np.int32))
    assert ulp.max() <= 1 and (ulp == 0).mean() >= 0.95, (ulp.max(), (ulp == 0).mean())
    assert rel_l2(cov, wcov) < 1e-5
    mu64, _ = fad.calc_embd_statistics(X.astype(np.float64))   # float64 in -> float64 mean, like np.mean
    assert mu64.dtype == np.float64 and rel_l2(mu64, X.astype(np.float64).mean(0)) < 1e-6
    # file-wise online merge (fadtk/utils.py:19-46) against one pass over all frames
    parts = np.array_split(X, 7)
    mu2, cov2 = fad.calculate_embd_statistics_online(parts)
    omu, ocov = ofad.embd_statistics_online(parts)    # the reference's arithmetic on the same fp16 files
    exact_mu, exact_cov = ofad.embd_statistics_online([p.astype(np.float64) for p in parts])
    assert rel_l2(mu2, exact_mu) < 1e-6 and rel_l2(cov2, exact_cov) < 1e-5
    # the reference's quirk, bounded: fp16 rounding of the per-file means (2^-11 relative per element) and what it does
    # to the cross terms of the Chan update (largest for few frames per file: 3e-4 at 6 frames, 1e-4 at 48)
    assert rel_l2(mu2, omu) < 2.0 ** -11 and rel_l2(cov2, ocov) < 5e-4, (rel_l2(mu2, omu), rel_l2(cov2, ocov))
    assert np.allclose(cov2, cov2.T)


def test_fad_vs_reference_functions():
    """the product against tests/golden/fad.npz: outputs of the reference's own calc_embd_statistics,
    calc_frechet_distance, calculate_embd_statistics_online and score_inf (fadtk/fad.py:41-47, 50-119, 303-350,
    fadtk/utils.py:13-46) run unmodified on seeded fp16 embeddings (tests/golden/make_fad_golden.py)."""
    import os
    from diffmusic_b200 import fad
    from tests.conftest import GOLDEN
    z = np.load(os.path.join(GOLDEN, "fad.npz"))
    for name, (n1, n2, d, parts) in stubs.FAD_CASES.items():
        a, b = stubs.fad_embeddings(name)
        mu1, c1 = fad.calc_embd_statistics(a)
        mu2, c2 = fad.calc_embd_statistics(b)
        for got, want in ((mu1, z[name + "_mu1"]), (mu2, z[name + "_mu2"])):
            assert got.dtype == np.float16
            ulp = np.abs(got.view(np.int16).astype(np.int32) - want.view(np.int16).astype(np.int32))
            assert ulp.max() <= 1 and (ulp == 0).mean() >= 0.95
        assert rel_l2(c1, z[name + "_cov1"]) < 1e-5 and rel_l2(c2, z[name + "_cov2"]) < 1e-5
        # distance from the reference's own statistics (isolates the solver) and from the product's (end to end)
        want = float(z[name + "_fd"])
        got = fad.calc_frechet_distance(z[name + "_mu1"], z[name + "_cov1"], z[name + "_mu2"], z[name + "_cov2"])
        assert abs(got - want) <= 1e-8 * want, (name, got, want)
        got2 = fad.calc_frechet_distance(mu1, c1, mu2, c2)
        assert abs(got2 - want) <= 1e-4 * want, (name, got2, want)
        omu, ocov = fad.calculate_embd_statistics_online(np.array_split(a, parts))
        assert rel_l2(omu, z[name + "_online_mu"]) < 2.0 ** -11 and rel_l2(ocov, z[name + "_online_cov"]) < 5e-4
    a, b = stubs.fad_embeddings("d128")
    mu_b, cov_b = fad.calc_embd_statistics(b)
    np.random.seed(1234)
    r = fad.score_inf(mu_b, cov_b, a, steps=6, min_n=200)
    want = z["inf_points"]
    pts = np.array(r.points)
    assert np.array_equal(pts[:, 0], want[:, 0])                 # same sample sizes, same rows (numpy's global generator)
    assert np.allclose(pts[:, 1], want[:, 1], rtol=1e-4, atol=0)
    assert abs(r.score - float(z["inf_score"])) <= 1e-3 * abs(float(z["inf_score"]))   # intercept of a 6-point fit


# ------------------------------------------------------------------------------------------------ Frechet distance
def _gauss_stats(rng, n, d, scale, shift):
    x = (rng.standard_normal((n, d)) * scale + shift)
    return x.mean(0), np.cov(x, rowvar=False)


@pytest.mark.parametrize("d", [1, 2, 7, 64, 128, 257, 512])
def test_jacobi_eigenvalues_vs_numpy(d):
    """one-sided Jacobi row rotations (dm_sym_eig_jacobi): eigenvalues of a symmetric PSD matrix vs numpy eigvalsh."""
    from diffmusic_b200 import _lib
    rng = np.random.default_rng(d)
    a = rng.standard_normal((d + 3, d))
    c = a.T @ a / (d + 2)
    w = torch.from_numpy(c.copy()).to(DEV)
    eig = torch.empty(d, device=DEV, dtype=torch.float64)
    state = torch.zeros(8, device=DEV, dtype=torch.int64)
    _lib.call("dm_sym_eig_jacobi", w.data_ptr(), d, 30, 1e-14, eig.data_ptr(), state.data_ptr(), _lib.stream())
    got = np.sort(eig.cpu().numpy())
    want = np.sort(np.linalg.eigvalsh(c))
    assert np.abs(got - want).max() <= 1e-12 * want.max()
    assert 1 <= int(state[3]) < 30 or d == 1  # converged before the sweep cap
    g = w.cpu().numpy() @ w.cpu().numpy().T  # rows ended orthogonal
    off = g - np.diag(np.diag(g))
    assert np.abs(off).max() <= 1e-12 * np.abs(np.diag(g)).max()


@pytest.mark.parametrize("d,n1,n2", [(16, 100, 80), (128, 2000, 1500), (512, 3000, 2500), (768, 4000, 4000),
                                     (64, 40, 500), (128, 30, 20), (1024, 600, 2000)])
def test_frechet_distance_vs_reference_formula(d, n1, n2):
    """dm_frechet_distance vs the restated fadtk calc_frechet_distance (scipy eig of C1 C2), incl. rank-deficient
    covariances (n < d), where the reference itself only resolves tr sqrt(C1 C2) to ~1e-8 relative."""
    from diffmusic_b200 import fad
    from oracle import fad as ofad
    rng = np.random.default_rng(d + n1)
    mu1, c1 = _gauss_stats(rng, n1, d, rng.uniform(0.1, 2.0, d), rng.standard_normal(d))
    mu2, c2 = _gauss_stats(rng, n2, d, 1.3, 0.2)
    got = fad.calc_frechet_distance(mu1, c1, mu2, c2)
    want = float(np.real(ofad.calc_frechet_distance(mu1, c1, mu2, c2)))
    tol = 1e-9 if min(n1, n2) > d else 5e-7
    assert abs(got - want) <= tol * abs(want), (got, want)
    assert fad.calc_frechet_distance(mu1, c1, mu1, c1) <= 1e-9 * np.trace(c1)  # identical Gaussians -> 0
    with pytest.raises(AssertionError):
        fad.calc_frechet_distance(mu1, c1, mu2[:-1], c2[:-1, :-1])


def test_fad_inf_matches_reference_loop():
    """score_inf: same NumPy draws as the reference loop, statistics / distances on the GPU."""
    from diffmusic_b200 import fad
    from oracle import fad as ofad
    rng = np.random.default_rng(5)
    d, N = 128, 6000
    emb = (rng.standard_normal((N, d)) * rng.uniform(0.5, 1.5, d) + 0.1).astype(np.float16)
    mu_b, cov_b = _gauss_stats(rng, 5000, d, 1.0, 0.0)
    np.random.seed(123)
    want = ofad.score_inf(mu_b, cov_b, emb, steps=7, min_n=500)  # the same fp16 rows: fp16 means
    np.random.seed(123)
    got = fad.score_inf(mu_b, cov_b, emb, steps=7, min_n=500)
    assert [p[0] for p in got.points] == [p[0] for p in want[3]]
    for (n, a), (_, b) in zip(got.points, want[3]):
        # the covariance comes from the tcgen05 moment kernel (fp32 TMEM accumulation over 256-row slabs, summed in
        # float64): ~1e-6 relative per entry, see test_fad_statistics_vs_numpy
        assert abs(a - np.real(b)) <= 2e-6 * abs(b), n
    assert abs(got.score - want[0]) <= 1e-5 * abs(want[0]) and abs(got.slope - want[1]) <= 1e-3 * abs(want[1])
    assert abs(got.r2 - want[2]) <= 1e-4


def test_gather_rows_exact():
    from diffmusic_b200 import _lib
    for d in (128, 100):
        x = torch.randn(1000, d, device=DEV).half()
        idx = torch.randint(0, 1000, (3333,), device=DEV)
        out = torch.empty((3333, d), device=DEV, dtype=torch.float16)
        _lib.call("dm_fad_gather_rows", x.data_ptr(), 1000, d, idx.data_ptr(), 3333, out.data_ptr(), _lib.stream())
        assert torch.equal(out, x[idx])


# ------------------------------------------------------------------------------------------------ eval metrics
@pytest.mark.parametrize("hop,pad", [(160, "constant"), (512, "constant"), (512, "reflect"), (1024, "constant")])
def test_log_spectral_distance_vs_reference_formula(hop, pad):
    from diffmusic_b200 import metrics
    from oracle import metrics as om
    L = 16000 + 37
    bg = stubs.synth_clips(3, L).numpy()
    ev = (stubs.synth_clips(3, L, first=7) * 0.8).numpy()
    ev[1, 100] = np.nan
    ev[2, 5000] = np.inf
    ev[2, 5001] = -np.inf
    lsd = metrics.LogSpectralDistance(16000, 1024, hop, 1e-10, pad_mode=pad)
    want = om.lsd_score(bg, ev, 1024, hop, 1e-10, output_mean=False, pad_mode=pad)
    got = lsd.score(bg, ev, output_mean=False)
    assert got.shape == want.shape == (3,)
    assert np.abs(got - want).max() <= 2e-5 * np.abs(want).max()
    assert abs(lsd.score(bg, ev) - float(om.lsd_score(bg, ev, 1024, hop, 1e-10, pad_mode=pad))) <= 2e-5 * want.mean()
    assert lsd.score(bg, bg) == 0.0
    with pytest.raises(NotImplementedError):
        metrics.LogSpectralDistance(n_fft=2048)


@pytest.mark.parametrize("name", ["b1_t26", "b3_t41_own_phase", "b2_t9"])
def test_mel_to_waveform_with_phase_vs_reference(name):
    """mel_spectrogram_to_waveform_with_phase through dm_istft_mel_phase against the output of the reference function
    itself (tests/golden/istft.npz: as-is, clipped and zero-padded lengths; phase shared by the batch or per clip) and
    the float64 oracle."""
    import os
    from diffmusic_b200.istft import mel_spectrogram_to_waveform_with_phase
    from oracle import istft as oi
    from tests.conftest import GOLDEN
    want = np.load(os.path.join(GOLDEN, "istft.npz"))[name]
    mel, phase = stubs.istft_inputs(name)
    length = stubs.ISTFT_CASES[name][4]
    got = mel_spectrogram_to_waveform_with_phase(mel.to(DEV), phase.to(DEV), original_waveform_length=length)
    assert got.dtype == torch.float32 and tuple(got.shape) == want.shape
    assert rel_l2(got, want) < 3e-6       # fp32 on both sides
    ref64 = oi.mel_spectrogram_to_waveform_with_phase(mel.numpy(), phase.numpy(), original_waveform_length=length)
    assert rel_l2(got, ref64) < 2e-6
    n = 160 * (stubs.ISTFT_CASES[name][1] - 1)
    assert not got[:, n:].any()
    # a (B, T, 64) mel without the channel axis and fp16 inputs (the pipeline runs in fp16) go through the same call
    again = mel_spectrogram_to_waveform_with_phase(mel[:, 0].to(DEV), phase.to(DEV), original_waveform_length=length)
    assert torch.equal(again, got)
    half = mel_spectrogram_to_waveform_with_phase(mel.half().to(DEV), phase.to(DEV), original_waveform_length=length)
    up = mel_spectrogram_to_waveform_with_phase(mel.half().float().to(DEV), phase.to(DEV),
                                                original_waveform_length=length)
    assert torch.equal(half, up)


def test_mel_to_waveform_with_phase_full_size():
    """BASELINE size (16 clips, 1001 frames -> 160 000 samples): two clips against the float64 oracle, every clip through
    the positive homogeneity of the chain (relu(W (2 mel)) = 2 relu(W mel), and scaling by a power of two commutes with
    every rounding, so doubling the mel doubles the waveform bit for bit); other hops and a two-frame input go against
    the oracle."""
    import math
    from diffmusic_b200.istft import mel_spectrogram_to_waveform_with_phase
    from oracle import istft as oi
    B, T = 16, 1001
    g = torch.Generator().manual_seed(21)
    mel = torch.rand(B, 1, T, 64, generator=g) * 6.0 - 1.0
    phase = (torch.rand(1, 513, T, generator=g) * 2.0 - 1.0) * math.pi
    got = mel_spectrogram_to_waveform_with_phase(mel.to(DEV), phase.to(DEV), original_waveform_length=160000)
    assert tuple(got.shape) == (B, 160000)
    for b in (0, 15):
        want = oi.mel_spectrogram_to_waveform_with_phase(mel[b:b + 1].numpy(), phase.numpy(),
                                                         original_waveform_length=160000)
        assert rel_l2(got[b:b + 1], want) < 2e-6
    twice = mel_spectrogram_to_waveform_with_phase((2.0 * mel).to(DEV), phase.to(DEV), original_waveform_length=160000)
    assert torch.equal(twice, 2.0 * got)
    for hop, T2 in ((512, 37), (160, 2), (1024, 5)):
        m2 = torch.rand(2, 1, T2, 64, generator=g) * 4.0
        p2 = (torch.rand(2, 513, T2, generator=g) * 2.0 - 1.0) * math.pi
        y = mel_spectrogram_to_waveform_with_phase(m2.to(DEV), p2.to(DEV), hop_length=hop)
        want = oi.mel_spectrogram_to_waveform_with_phase(m2.numpy(), p2.numpy(), hop_length=hop)
        assert tuple(y.shape) == want.shape == (2, hop * (T2 - 1))
        assert rel_l2(y, want) < 2e-6
    with pytest.raises(NotImplementedError):
        mel_spectrogram_to_waveform_with_phase(mel.to(DEV), phase.to(DEV), n_fft=2048, win_length=2048)
    with pytest.raises(RuntimeError):
        mel_spectrogram_to_waveform_with_phase(mel.to(DEV), phase[..., :1000].to(DEV))


def test_waveform_to_spectrogram_vs_reference():
    """waveform_to_spectrogram (diffmusic/utils.py:11-20) through dm_stft_spectrogram: the reference function's own output
    (tests/golden/istft.npz), the float64 oracle at full size, and the magnitude against PhaseRetrievalOperator.forward
    (bit-equal on the same 64-thread frame-pair engine, fp32 rounding apart from the warp-per-pair engine)."""
    import os
    from oracle import istft as oi
    from tests.conftest import GOLDEN
    from tests.test_oracle_vs_golden import _spectrogram_error
    z = np.load(os.path.join(GOLDEN, "istft.npz"))
    for name, L in stubs.SPECTROGRAM_CASES.items():
        wav = stubs.synth_clips(2, L, first=70)
        mag, phase = dm.waveform_to_spectrogram(wav.to(DEV))
        assert tuple(mag.shape) == tuple(phase.shape) == z[name + "_mag"].shape
        rel, dphi = _spectrogram_error(mag.cpu().numpy(), phase.cpu().numpy(), z[name + "_mag"], z[name + "_phase"])
        assert rel < 3e-6 and dphi < 2e-3
        m1, p1 = dm.waveform_to_spectrogram(wav[1].to(DEV))          # 1-D input, like torch.stft
        assert torch.equal(m1, mag[1]) and torch.equal(p1, phase[1])
    L = 160000
    wav = stubs.synth_clips(3, L, first=80)
    mag, phase = dm.waveform_to_spectrogram(wav.to(DEV))
    assert tuple(mag.shape) == (3, 513, 1001)
    wm, wp = oi.waveform_to_spectrogram(wav[2:3].numpy())
    rel, dphi = _spectrogram_error(mag[2:3].cpu().numpy(), phase[2:3].cpu().numpy(), wm, wp)
    assert rel < 3e-6 and dphi < 2e-3
    op = dm.PhaseRetrievalOperator(1024, 160, 1024, noiser=dm.get_noiser("gaussian", 0.0))
    from diffmusic_b200 import _lib
    assert rel_l2(op.forward(wav.to(DEV)), mag) < 1e-6
    _lib.call("dm_stft_set_engine", 2)
    try:
        assert torch.equal(op.forward(wav.to(DEV)), mag)
    finally:
        _lib.call("dm_stft_set_engine", 0)
    ph = phase.cpu().numpy()
    assert np.isin(ph[:, [0, 512]], np.float32([0.0, np.pi])).all()   # real DC / Nyquist bins: angle 0 or pi exactly
    m512, p512 = dm.waveform_to_spectrogram(wav[:1, :20000].to(DEV), hop_length=512)
    wm, wp = oi.waveform_to_spectrogram(wav[:1, :20000].numpy(), hop_length=512)
    rel, dphi = _spectrogram_error(m512.cpu().numpy(), p512.cpu().numpy(), wm, wp)
    assert tuple(m512.shape) == (1, 513, 40) and rel < 3e-6 and dphi < 2e-3
    with pytest.raises(RuntimeError):
        dm.waveform_to_spectrogram(wav[:, :500].to(DEV))


def test_mean_squared_error_vs_reference_formula():
    from diffmusic_b200 import metrics
    from oracle import metrics as om
    bg = stubs.synth_clips(4, 40001).numpy()
    ev = stubs.synth_clips(4, 40001, first=9).numpy()
    ev[0, 3] = np.nan
    bg[1, 7] = np.inf
    for red in ("mean", "sum"):
        got = metrics.MeanSquaredError(red).score(bg, ev)
        want = float(om.mse_score(bg, ev, red))
        assert abs(got - want) <= 1e-6 * abs(want)
    ragged_bg = [bg[0], bg[1][:30000]]
    ragged_ev = [ev[0][:25000], ev[1]]
    got = metrics.MeanSquaredError("sum").score(ragged_bg, ragged_ev)
    want = sum(float(np.mean((np.nan_to_num(a[:min(len(a), len(b))], nan=0, posinf=1, neginf=-1)
                              - np.nan_to_num(b[:min(len(a), len(b))], nan=0, posinf=1, neginf=-1)) ** 2))
               for a, b in zip(ragged_bg, ragged_ev))
    assert abs(got - want) <= 1e-6 * abs(want)
    with pytest.raises(AssertionError):
        metrics.MeanSquaredError("median")


# ------------------------------------------------------------------------------------------------ CUDA-graph replay
@pytest.mark.parametrize("sched_name,op_name,eta", [("ddim", "inpainting", 0.0), ("dps", "super_resolution", 0.0),
                                                    ("dps", "inpainting", 0.5), ("mpgd", "inpainting", 1.0),
                                                    ("dsg", "phase_retrieval", 1.0),
                                                    ("diffmusic", "dereverberation", 1.0)])
def test_graphed_step_equals_eager_step(sched_name, op_name, eta):
    """one captured graph replayed over several timesteps == the eager `.step` (same generators, same IR draws)."""
    B = 2
    vae, voc = stubs.StubVAE().to(DEV), stubs.StubVocoder().to(DEV)
    x, e = stubs.synth_latents(B, 25)
    x, e = x.to(DEV), e.to(DEV)
    op = _ops()[op_name]
    sched = dm.get_scheduler(sched_name)(operator=op, **stubs.MUSICLDM_SCHED)
    sched.set_timesteps(500)
    torch.manual_seed(321)
    meas = op.forward(stubs.synth_clips(1, L1, first=50).to(DEV))
    kw = dict(eta=eta, measurement=meas, vae=vae, vocoder=voc, original_waveform_length=L1)
    if RATES[sched_name] is not None:
        kw.update(ip_guidance_rate=RATES[sched_name], supervised_space="mel_spectrogram")
    graphed = dm.GraphedGuidedStep(sched, tuple(x.shape), **kw)
    g_eager, g_graph = stubs.step_generators(B), stubs.step_generators(B)
    xe, xg = x.clone(), x.clone()
    for t in (999, 501, 1):
        torch.manual_seed(40 + t)  # dereverb IR (global CPU generator)
        a = sched.step(e, t, xe, generator=g_eager, **kw)
        torch.manual_seed(40 + t)
        b = graphed(e, t, xg, generator=g_graph)
        assert rel_l2(b.prev_sample, a.prev_sample) < 1e-6, t
        assert rel_l2(b.pred_original_sample, a.pred_original_sample) < 1e-6, t
        assert abs(float(b.loss.float().ravel()[0]) - float(a.loss.float().ravel()[0])) <= 1e-6 * abs(float(a.loss.float().ravel()[0])) + 1e-12
        xe, xg = a.prev_sample, b.prev_sample  # chain the trajectory


@pytest.mark.parametrize("two_graphs", [False, True])
def test_host_pipelined_steps_equal_direct_replays(two_graphs):
    """HostPipelinedStep (pinned-host latents, uploads / downloads overlapped on side streams) returns exactly what the
    same graph gives for device-resident inputs, for a run long enough to reuse both staging slots several times."""
    B, n = 3, 7
    vae, voc = stubs.StubVAE().to(DEV), stubs.StubVocoder().to(DEV)
    op = _ops()["super_resolution"]
    sched = dm.get_scheduler("dps")(operator=op, **stubs.MUSICLDM_SCHED)
    sched.set_timesteps(500)
    meas = op.forward(stubs.synth_clips(1, L1, first=50).to(DEV))
    kw = dict(eta=0.0, measurement=meas, vae=vae, vocoder=voc, original_waveform_length=L1, ip_guidance_rate=5e-4,
              supervised_space="mel_spectrogram")
    graphed = dm.GraphedGuidedStep(sched, (B, 8, 25, 16), **kw)
    lat = [stubs.synth_latents(B, 25, first=10 * i) for i in range(n)]
    ts = [int(t) for t in sched.timesteps[:n]]
    want = []
    for (x, e), t in zip(lat, ts):
        o = graphed(e.to(DEV), t, x.to(DEV))
        want.append((o.prev_sample.cpu(), o.loss_per_clip.reshape(-1).cpu()))
    xs = [x.pin_memory() for x, _ in lat]
    es = [e.pin_memory() for _, e in lat]
    prev = [torch.empty_like(xs[0]).pin_memory() for _ in range(n)]
    loss = [torch.empty(B).pin_memory() for _ in range(n)]
    second = dm.GraphedGuidedStep(sched, (B, 8, 25, 16), **kw) if two_graphs else None
    pipe = dm.HostPipelinedStep(graphed, second)
    pipe.prefetch(es[0], xs[0])
    for i in range(n):
        if i + 1 < n:
            pipe.prefetch(es[i + 1], xs[i + 1])
        pipe.step(ts[i], prev[i], loss[i])
    with pytest.raises(RuntimeError):
        pipe.step(ts[0], prev[0], loss[0])  # nothing prefetched
    pipe.drain()
    torch.cuda.synchronize()
    for i in range(n):
        assert torch.equal(prev[i], want[i][0]), i
        assert torch.equal(loss[i], want[i][1]), i


@pytest.mark.parametrize("nf", [6, 7, 8, 13, 16, 22])
@pytest.mark.parametrize("op_name", ["inpainting", "phase_retrieval"])
def test_stft_kernels_agree_for_every_tile_size(op_name, nf, monkeypatch):
    """warp-per-frame-pair kernel (default for <= 16 frames per tile) vs 64-thread frame-pair kernel vs frame-at-a-time
    kernel, over tile sizes with full, partial and odd frame counts: loss and gradient agree to fp32 rounding, in both
    supervised spaces."""
    from diffmusic_b200 import _lib
    monkeypatch.setenv("DM_STFT_FRAMES_PER_TILE", str(nf))
    B, L = 3, L1  # T = 101 frames: every tile size leaves a partial last tile, some with an odd frame count
    wav = stubs.synth_clips(B, L).to(DEV)
    op = _inpaint() if op_name == "inpainting" else dm.PhaseRetrievalOperator(noiser=_noiser())
    meas = op.forward(stubs.synth_clips(1, L, first=50).to(DEV))
    engines = (("auto", 0), ("pair", 2), ("frame", 1))
    for space in ("mel_spectrogram", "wav_form"):
        res = {}
        for name, eng in engines:
            _lib.call("dm_stft_set_engine", eng)
            try:
                res[name] = _loss_grad(op, wav, meas, space)
            finally:
                _lib.call("dm_stft_set_engine", 0)
        for name, tol in (("auto", 5e-6), ("pair", 2e-6)):  # 32 x 32 transform vs radix-8 passes: fp32 rounding apart
            assert rel_l2(res[name][0], res["frame"][0]) < 1e-6, (space, name)
            assert rel_l2(res[name][1], res["frame"][1]) < tol, (space, name)
    t = {}
    for name, eng in engines:
        _lib.call("dm_stft_set_engine", eng)
        try:
            t[name] = op.transform(op.forward(wav))
        finally:
            _lib.call("dm_stft_set_engine", 0)
    assert rel_l2(t["auto"], t["frame"]) < 1e-6
    assert rel_l2(t["pair"], t["frame"]) < 1e-6


@pytest.mark.parametrize("B,L", [(1, 16000), (3, 16000), (5, 32008), (2, 20016), (16, 160000), (3, 4104)])
def test_stream_resampling_kernels_are_bit_identical(B, L):
    """the persistent-grid scale-2 resampling kernels (cp.async-staged swizzled windows forward, contiguous ranges with
    the scales of the touched clips adjoint, 256-bit stores; dm_set_tuning DM_TUNE_STREAM_KERNELS = 1, the default)
    against the one-CTA-per-2048-samples kernels: A(x), loss and gradient bit-identical in both supervised spaces, for
    one clip, clips that straddle CTA ranges, short rows and the BASELINE batch."""
    from diffmusic_b200 import _lib
    op = dm.SuperResolutionOperator(16000, scale=2, noiser=_noiser())
    wav = stubs.synth_clips(B, L).to(DEV)
    meas = op.forward(stubs.synth_clips(1, L, first=50).to(DEV))
    res = {}
    for knob in (2, 1, 0):  # 2: persistent kernels whatever the batch; 1: the default choice; 0: plain kernels
        _lib.call("dm_set_tuning", 0, knob)
        try:
            res[knob] = (op.forward(wav),) + tuple(op.fused_loss_and_grad(wav, meas, "mel_spectrogram")) + tuple(
                op.fused_loss_and_grad(wav, meas, "wav_form"))
        finally:
            _lib.call("dm_set_tuning", 0, 1)
    for knob in (2, 1):
        for a, b in zip(res[knob], res[0]):
            assert torch.equal(a, b)


@pytest.mark.parametrize("nf", [6, 10, 14])
@pytest.mark.parametrize("B,L", [(3, 16000), (2, 20011), (1, 4099), (2, 160000)])
def test_fused_resampling_chain_is_bit_identical(B, L, nf, monkeypatch):
    """dm_stft_guidance_fir2 (the scale-2 sinc resampling computed inside the STFT kernel, its spans never written to
    HBM; opt-in, DM_STFT_FUSE_FIR=1) against dm_resample_fwd -> dm_stft_guidance with the same tile size: same arithmetic
    per sample, so loss and gradient are bit-identical; odd lengths, lengths whose last tile is a single frame (L = 16000, nf = 10: its mirrored
    right edge starts one sample outside the tile's span), and the 10 s clip; and against the oracle."""
    from diffmusic_b200 import _lib
    monkeypatch.setenv("DM_STFT_FRAMES_PER_TILE", str(nf))
    op = dm.SuperResolutionOperator(16000, scale=2, noiser=_noiser())
    buf = torch.zeros((B, (L + 3) // 4 * 4), device=DEV)  # rows the kernel can read with 128-bit loads (the vocoder
    wav = buf[:, :L]                                      # output of the pipelines: a [:, :L] slice of aligned rows)
    wav.copy_(stubs.synth_clips(B, L))
    meas = op.forward(stubs.synth_clips(1, L, first=50).to(DEV))
    op._ref_mel(meas)  # cached from here on: the counted launches below are the chain's own
    res = {}
    for fuse in ("1", "0"):
        monkeypatch.setenv("DM_STFT_FUSE_FIR", fuse)
        n0 = _lib.launch_count()
        res[fuse] = op.fused_loss_and_grad(wav, meas, "mel_spectrogram")
        res[fuse + "n"] = _lib.launch_count() - n0
    assert res["1n"] == 2 and res["0n"] == 3  # the resampling launch is gone
    if ((L + 1) // 2) % 4 == 0:  # otherwise the unfused chain resamples with the staged polyphase kernel (other rounding)
        assert torch.equal(res["1"][0], res["0"][0])
        assert torch.equal(res["1"][1], res["0"][1])
    else:
        assert rel_l2(res["1"][0], res["0"][0]) < 1e-6 and rel_l2(res["1"][1], res["0"][1]) < TOL
    if L <= 20011:
        oop = oo.OracleOperator("super_resolution", scale=2)
        for i in range(B):
            w = wav[i:i + 1].cpu().clone().requires_grad_(True)
            want = torch.linalg.norm(oop.transform(meas.cpu()) - oop.transform(oop.forward(w)))
            (gw,) = torch.autograd.grad(want, w)
            assert abs(float(res["1"][0][i]) - float(want)) < TOL * float(want)
            assert rel_l2(res["1"][1][i:i + 1], gw) < TOL


# ------------------------------------------------------------------------------------------------ update kernels, all paths
@pytest.mark.parametrize("n_clip", [3200, 32000, 38400, 3203])
@pytest.mark.parametrize("kind", ["dsg", "diffmusic"])
def test_norm_update_kernel_paths(kind, n_clip):
    """dm_sched_{dsg,diffmusic}_update against the reference algebra (scheduling_dsg.py:189-224,
    scheduling_diffmusic.py:59-68,191-223) for the register-cached path (<= 32768 elements per clip), the re-reading
    path (longer clips) and the scalar path (n_clip % 4 != 0); 3 clips, per-clip norms."""
    from diffmusic_b200 import _lib
    B = 3
    g = torch.Generator().manual_seed(n_clip)
    x0, ep, g0, z = (torch.randn(B, n_clip, generator=g) for _ in range(4))
    g0 = g0 * 37.0
    sa, sp, dirc, std, rate, e = 0.7311, 0.8123, 0.5377, 0.2214, 0.08, 1e-8
    r = float(torch.sqrt(torch.tensor(n_clip)) * std)
    want = []
    for b in range(B):
        gb = g0[b] * (1.0 / 1000.0) / sa
        mean = sp * x0[b] + dirc * ep[b]
        gn = torch.linalg.norm(gb)
        if kind == "dsg":
            d_star = -r * gb / (gn + e)
            d_s = std * z[b]
            mix = d_s + rate * (d_star - d_s)
            want.append(mean + r * mix / (torch.linalg.norm(mix) + e))
        else:
            gt = gb / (gn + e) * torch.linalg.norm(z[b])
            want.append(mean + std * osteps.slerp(z[b], -gt, rate))
    want = torch.stack(want)
    d = [t.to(DEV).contiguous() for t in (x0, ep, g0, z)]
    prev = torch.empty_like(d[0])
    if kind == "dsg":
        _lib.call("dm_sched_dsg_update", d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(),
                  prev.data_ptr(), B, n_clip, sa, sp, dirc, std, rate, r, 1.0 / 1000.0, e, None, _lib.stream())
    else:
        _lib.call("dm_sched_diffmusic_update", d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(),
                  prev.data_ptr(), B, n_clip, sa, sp, dirc, std, rate, 1.0 / 1000.0, e, 0.9995, None, _lib.stream())
    assert rel_l2(prev, want) < 2e-6
    for b in range(B):
        assert rel_l2(prev[b], want[b]) < 2e-6


def test_slerp_linear_branch_on_device():
    """|cos| > 0.9995 (z and -g nearly parallel): the kernel takes the linear-interpolation branch per clip."""
    from diffmusic_b200 import _lib
    n = 3200
    g = torch.Generator().manual_seed(3)
    z = torch.randn(2, n, generator=g)
    g0 = torch.stack([-z[0] * 5.0 + 1e-4 * torch.randn(n, generator=g), torch.randn(n, generator=g)])  # clip 0 parallel
    x0, ep = torch.randn(2, n, generator=g), torch.randn(2, n, generator=g)
    sa, sp, dirc, std, rate, e = 0.9, 0.8, 0.5, 0.3, 0.08, 1e-8
    want = []
    for b in range(2):
        gb = g0[b] * 1e-3 / sa
        gt = gb / (torch.linalg.norm(gb) + e) * torch.linalg.norm(z[b])
        want.append(sp * x0[b] + dirc * ep[b] + std * osteps.slerp(z[b], -gt, rate))
    d = [t.to(DEV).contiguous() for t in (x0, ep, g0, z)]
    prev = torch.empty_like(d[0])
    _lib.call("dm_sched_diffmusic_update", d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(),
              prev.data_ptr(), 2, n, sa, sp, dirc, std, rate, 1e-3, e, 0.9995, None, _lib.stream())
    assert rel_l2(prev, torch.stack(want)) < 2e-6


# ------------------------------------------------------------------------------------------------ edge cases
@pytest.mark.parametrize("L", [513, 1000, 1023, 16001, 40000])
def test_ragged_lengths_against_oracle(L):
    """shortest legal clip (reflect padding needs L > 512), fewer frames than one tile, L not a multiple of the hop."""
    wav = stubs.synth_clips(2, L).to(DEV)
    ref = stubs.synth_clips(1, L, first=50)
    op, oop = dm.IdentityOperator(16000), oo.OracleOperator("identity")
    assert rel_l2(op.transform(wav), oop.transform(wav.cpu())) < TOL
    loss, g = _loss_grad(op, wav, ref.to(DEV), "mel_spectrogram")
    for i in range(2):
        w = wav[i:i + 1].cpu().clone().requires_grad_(True)
        want = torch.linalg.norm(oop.transform(ref) - oop.transform(w))
        (gw,) = torch.autograd.grad(want, w)
        assert abs(float(loss[i]) - float(want)) < TOL * float(want)
        assert rel_l2(g[i:i + 1], gw) < TOL
    sr, osr = dm.SuperResolutionOperator(16000, 2, _noiser()), oo.OracleOperator("super_resolution", scale=2)
    assert rel_l2(sr.forward(wav), osr.forward(wav.cpu())) < TOL
    if L >= 1000:
        dv = dm.MusicDereverberationOperator(800, 0.85, _noiser())
        torch.manual_seed(L)
        y = dv.forward(wav)
        assert rel_l2(y, oo.a_dereverb(wav.cpu(), dv.last_ir)) < TOL


def test_shape_errors_are_python_exceptions():
    inp = _inpaint()
    with pytest.raises(ValueError):
        inp.forward(torch.zeros(1, 12345, device=DEV))          # mask / data length mismatch
    op = dm.IdentityOperator(16000)
    with pytest.raises(ValueError):
        op.guidance_loss(torch.zeros(1, 16000, device=DEV), torch.zeros(1, 8000, device=DEV), "mel_spectrogram")
    with pytest.raises(ValueError):
        op.guidance_loss(torch.zeros(1, 16000, device=DEV), torch.zeros(1, 16000, device=DEV), "nope")
    with pytest.raises(Exception):
        op.transform(torch.zeros(1, 300, device=DEV))             # shorter than the reflect padding
    with pytest.raises(NotImplementedError):
        dm.PhaseRetrievalOperator(n_fft=2048, hop_length=512, win_length=2048)
    with pytest.raises(NotImplementedError):
        dm.MusicDereverberationOperator(ir_length=9000)


def test_large_batch_matches_small_batches():
    """B = 48 in one launch == the same clips in batches of 1 (no cross-clip coupling anywhere)."""
    L, B = 16000, 48
    wav = stubs.synth_clips(B, L).to(DEV)
    meas = stubs.synth_clips(1, L, first=50).to(DEV)
    op = dm.SuperResolutionOperator(16000, 2, _noiser())
    m = op.forward(meas)
    loss, g = _loss_grad(op, wav, m, "mel_spectrogram")
    for i in (0, 17, 47):
        li, gi = _loss_grad(op, wav[i:i + 1], m, "mel_spectrogram")
        # tile boundaries (hence the fp32 summation grouping) depend on the launch shape: equal to rounding, not bitwise
        assert rel_l2(li, loss[i:i + 1]) < 1e-6 and rel_l2(gi, g[i:i + 1]) < 1e-5


# ------------------------------------------------------------------------------------------------ remaining section-8 rows
def test_generic_ratio_resample_kernels():
    """non-integer ratio (24 kHz -> 16 kHz: orig 3, new 2) goes through the generic FIR kernels (a3)."""
    import torchaudio
    from diffmusic_b200 import _lib, tables
    kern, width, orig, new = tables.sinc_resample_kernel(24000, 16000)
    assert (orig, new) == (3, 2)
    L, B = 7001, 2
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, L, generator=g)
    rs = torchaudio.transforms.Resample(24000, 16000)
    xx = x.clone().requires_grad_(True)
    y = rs(xx)
    Ly = y.shape[1]
    xd, kd = x.to(DEV).contiguous(), kern.to(DEV).contiguous()
    yd = torch.empty(B, Ly, device=DEV)
    _lib.call("dm_resample_fwd", xd.data_ptr(), xd.stride(0), L, B, kd.data_ptr(), kd.shape[0], kd.shape[1], orig,
              width, yd.data_ptr(), Ly, _lib.stream())
    assert rel_l2(yd, y.detach()) < 1e-5
    yb = torch.randn(B, Ly, generator=g)
    (gx,) = torch.autograd.grad((y * yb).sum(), xx)
    ybd = yb.to(DEV).contiguous()
    partial = torch.ones(B, 1, device=DEV)          # loss = 1 -> scale 1
    dwav, loss = torch.empty(B, L, device=DEV), torch.empty(B, device=DEV)
    _lib.call("dm_resample_adjoint", ybd.data_ptr(), 0, Ly, B, partial.data_ptr(), 1, kd.data_ptr(), kd.shape[0],
              kd.shape[1], orig, width, dwav.data_ptr(), dwav.stride(0), L, loss.data_ptr(), _lib.stream())
    assert rel_l2(dwav, gx) < 1e-5 and torch.allclose(loss.cpu(), torch.ones(B))


def test_style_guidance_mpgd_generic_path():
    """BASELINE config 5: StyleGuidanceOperator (forward = identity, transform = user torch callable) + MPGD: the
    residual comes from torch (no fused kernel), x0 and the MPGD update from the CUDA kernels."""
    class FakeClap:
        def get_gram_matrix(self, audio):           # (B, L) -> (B, 64, 64) gram of 64 strided feature maps
            f = audio.reshape(audio.shape[0], 64, -1)
            return f @ f.transpose(1, 2) / f.shape[-1]

    B = 2
    vae, voc = stubs.StubVAE(), stubs.StubVocoder()
    x, e = stubs.synth_latents(B, 25)
    ref = stubs.synth_clips(B, L1, first=50)
    op = dm.StyleGuidanceOperator(FakeClap())
    meas = op.forward(ref)

    class OracleStyle:
        forward = staticmethod(lambda d, **k: d)
        transform = staticmethod(lambda a: FakeClap().get_gram_matrix(a.float()))
        inverse_transform = staticmethod(lambda m, v: v(m.squeeze(1) if m.dim() == 4 else m))

    base = osteps.make_base(**stubs.MUSICLDM_SCHED)
    base.set_timesteps(500)
    want = osteps.per_clip_step("mpgd", base, OracleStyle, e, 501, x, measurement=meas, eta=0.0,
                                ip_guidance_rate=0.005, vae=vae, vocoder=voc, original_waveform_length=L1,
                                supervised_space="mel_spectrogram")
    sched = dm.MPGDScheduler(operator=op, **stubs.MUSICLDM_SCHED)
    sched.set_timesteps(500)
    got = sched.step(e.to(DEV), 501, x.to(DEV), eta=0.0, measurement=meas.to(DEV), vae=vae.to(DEV),
                     vocoder=voc.to(DEV), original_waveform_length=L1, ip_guidance_rate=0.005)
    assert rel_l2(got.prev_sample, want.prev_sample) < TOL
    assert rel_l2(got.pred_original_sample, want.pred_original_sample) < TOL
    assert rel_l2(got.loss_per_clip, want.loss) < TOL


def test_noisy_operator_inside_guidance():
    """GaussianNoise with sigma > 0 attached to an operator (noise.py:13-18): the noise is drawn by torch on the
    tensor's device exactly where the reference draws it, and added by dm_add_scaled; the VJP is unaffected."""
    sigma = 0.05
    wav = stubs.synth_clips(1, L1).to(DEV)
    ref = stubs.synth_clips(1, L1, first=50).to(DEV)
    op = dm.SuperResolutionOperator(16000, 2, dm.get_noiser("gaussian", sigma))
    clean = dm.SuperResolutionOperator(16000, 2, _noiser())
    torch.manual_seed(11)
    y = op.forward(wav)
    torch.manual_seed(11)
    n = torch.randn_like(clean.forward(wav))
    assert rel_l2(y, clean.forward(wav) + sigma * n) < 1e-6
    meas = clean.forward(ref)
    torch.manual_seed(12)
    loss, g = _loss_grad(op, wav, meas, "mel_spectrogram")
    # same draw, evaluated by the CPU oracle on the noisy prediction
    torch.manual_seed(12)
    n = torch.randn_like(meas)
    w = wav.cpu().clone().requires_grad_(True)
    oop = oo.OracleOperator("super_resolution", scale=2)
    pred = oop.forward(w) + sigma * n.cpu()
    want = torch.linalg.norm(oop.transform(meas.cpu()) - oop.transform(pred))
    (gw,) = torch.autograd.grad(want, w)
    assert abs(float(loss) - float(want)) < TOL * float(want)
    assert rel_l2(g, gw) < TOL


def test_fp16_pipeline_dtype_is_accepted():
    """run.py:218 loads the pipelines in fp16: latents, noise prediction and the networks are half precision.  The
    scheduler computes in fp32 (operators up-cast with .float(), operator.py:154,205,248) and returns the caller's
    dtype; the result stays close to the fp32 step."""
    B = 2
    vae, voc = stubs.StubVAE().to(DEV), stubs.StubVocoder().to(DEV)
    x, e = stubs.synth_latents(B, 25)
    op = dm.SuperResolutionOperator(16000, 2, _noiser())
    meas = op.forward(stubs.synth_clips(1, L1, first=50).to(DEV))
    sched = dm.DPSScheduler(operator=op, **stubs.MUSICLDM_SCHED)
    sched.set_timesteps(500)
    kw = dict(measurement=meas, original_waveform_length=L1, ip_guidance_rate=5e-4)
    full = sched.step(e.to(DEV), 501, x.to(DEV), vae=vae, vocoder=voc, **kw)
    half = sched.step(e.to(DEV).half(), 501, x.to(DEV).half(), vae=vae.half(), vocoder=voc.half(), **kw)
    assert half.prev_sample.dtype == torch.float16 and half.pred_original_sample.dtype == torch.float16
    assert torch.isfinite(half.prev_sample).all() and torch.isfinite(half.loss)
    assert rel_l2(half.prev_sample.float(), full.prev_sample) < 5e-3
    assert abs(float(half.loss) - float(full.loss)) < 2e-2 * float(full.loss)


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("n_lat", [25, 26])
def test_latent_io_dtypes_are_exact_roundings_of_the_fp32_step(dt, n_lat):
    """dm_sched_*_io read / write 16-bit latents directly: with no network in the loop (DDIM) the result must be the
    fp32 step on the up-cast inputs, rounded once to the latent dtype -- bit for bit (vector and scalar paths)."""
    B = 3
    x, e = stubs.synth_latents(B, n_lat)
    if n_lat == 26:  # odd element count per tensor -> scalar (W = 1) kernels
        x, e = x[..., :15].contiguous(), e[..., :15].contiguous()
    x, e = x.to(DEV).to(dt), e.to(DEV).to(dt)
    sched = dm.DDIMScheduler(operator=None, **stubs.MUSICLDM_SCHED)
    sched.set_timesteps(500)
    for t in (999, 501, 1):
        got = sched.step(e, t, x)
        want = sched.step(e.float(), t, x.float())
        assert got.prev_sample.dtype == dt and got.pred_original_sample.dtype == dt
        assert torch.equal(got.prev_sample, want.prev_sample.to(dt)), t
        assert torch.equal(got.pred_original_sample, want.pred_original_sample.to(dt)), t


@pytest.mark.parametrize("dt,tol", [(torch.float16, 5e-3), (torch.bfloat16, 4e-2)])
@pytest.mark.parametrize("sched_name,op_name,eta", [("dps", "super_resolution", 0.0), ("mpgd", "inpainting", 1.0),
                                                    ("dsg", "phase_retrieval", 1.0),
                                                    ("diffmusic", "inpainting", 1.0)])
def test_guided_steps_in_16_bit_pipelines(sched_name, op_name, eta, dt, tol):
    """every guided scheduler with 16-bit latents and 16-bit networks (run.py:218) stays close to its fp32 step; the
    outputs come back in the pipeline's dtype, the per-clip loss in fp32."""
    B = 2
    x, e = stubs.synth_latents(B, 25)
    op = _ops()[op_name]
    sched = dm.get_scheduler(sched_name)(operator=op, **stubs.MUSICLDM_SCHED)
    sched.set_timesteps(500)
    meas = op.forward(stubs.synth_clips(1, L1, first=50).to(DEV))
    kw = dict(eta=eta, measurement=meas, original_waveform_length=L1, ip_guidance_rate=RATES[sched_name])
    z = torch.randn(B, 8, 25, 16, generator=torch.Generator().manual_seed(7)).to(DEV)
    vae, voc = stubs.StubVAE().to(DEV), stubs.StubVocoder().to(DEV)
    full = sched.step(e.to(DEV), 501, x.to(DEV), vae=vae, vocoder=voc, _noise=z, **kw)
    low = sched.step(e.to(DEV).to(dt), 501, x.to(DEV).to(dt), vae=stubs.StubVAE().to(DEV).to(dt),
                     vocoder=stubs.StubVocoder().to(DEV).to(dt), _noise=z, **kw)
    assert low.prev_sample.dtype == dt and low.pred_original_sample.dtype == dt
    assert low.loss_per_clip.dtype == torch.float32 and torch.isfinite(low.prev_sample).all()
    assert rel_l2(low.prev_sample.float(), full.prev_sample) < tol
    assert rel_l2(low.pred_original_sample.float(), full.pred_original_sample) < tol
    assert rel_l2(low.loss_per_clip, full.loss_per_clip) < 4 * tol


@pytest.mark.parametrize("sched_name,op_name,eta", [("dps", "super_resolution", 0.0), ("mpgd", "inpainting", 1.0),
                                                    ("dsg", "inpainting", 1.0), ("diffmusic", "inpainting", 1.0)])
def test_fp16_step_against_the_reference_arithmetic_run_in_fp16(sched_name, op_name, eta):
    """run.py:218 loads the pipelines with torch_dtype=float16, so the REFERENCE does the scheduler algebra in fp16 (the
    0-d fp32 scalars promote to the tensors' dtype).  The oracle is run exactly like that -- half latents,
    half stand-in networks, half step noise from the same seeded generator, CPU -- and the product gets the same half
    inputs on the GPU: its fp16 step (fp32 algebra, rounded once) stays within a few fp16 ulps of the reference-in-fp16.
    DPS / MPGD take a given `variance_noise`, so their fp32 oracle run sees the same noise and the product must also be
    at least as close to it as the reference-in-fp16 is; DSG / DiffMusic always draw their noise in the tensors' dtype
    (scheduling_dsg.py:215-220), and a half draw is a different sample than a float draw, so only the fp16 runs compare."""
    vae, voc = stubs.StubVAE(), stubs.StubVocoder()
    x, e = stubs.synth_latents(1, 25)
    oop = {"super_resolution": oo.OracleOperator("super_resolution", scale=2),
           "inpainting": oo.OracleOperator("inpainting", mask=oo.inpaint_mask(1, 16000, "box", 0.25, 0.5))}[op_name]
    op = _ops()[op_name]
    meas = oop.forward(stubs.synth_clips(1, L1, first=50))
    base = osteps.make_base(**stubs.MUSICLDM_SCHED)
    base.set_timesteps(500)
    z = torch.randn(1, 8, 25, 16, generator=torch.Generator().manual_seed(7))
    noise_kw = (lambda dt: dict(variance_noise=z.to(dt))) if sched_name in ("dps", "mpgd") else \
        (lambda dt: dict(generator=torch.Generator().manual_seed(3)))
    kw = dict(eta=eta, ip_guidance_rate=RATES[sched_name], measurement=meas, original_waveform_length=L1,
              supervised_space="mel_spectrogram")
    want32 = osteps.reference_step(sched_name, base, oop, e, 501, x, vae=vae, vocoder=voc, **noise_kw(torch.float32), **kw)
    want16 = osteps.reference_step(sched_name, base, oop, e.half(), 501, x.half(), vae=stubs.StubVAE().half(),
                                   vocoder=stubs.StubVocoder().half(), **noise_kw(torch.float16), **kw)
    assert want16.prev_sample.dtype == torch.float16
    sched = dm.get_scheduler(sched_name)(operator=op, **stubs.MUSICLDM_SCHED)
    sched.set_timesteps(500)
    gkw = dict(kw, measurement=meas.to(DEV))
    if sched_name in ("dps", "mpgd"):
        gkw["variance_noise"] = z.to(DEV).half()
    else:
        gkw["generator"] = torch.Generator().manual_seed(3)
    got = sched.step(e.to(DEV).half(), 501, x.to(DEV).half(), vae=stubs.StubVAE().to(DEV).half(),
                     vocoder=stubs.StubVocoder().to(DEV).half(), **gkw)
    assert got.prev_sample.dtype == torch.float16
    err16 = rel_l2(got.prev_sample.float(), want16.prev_sample.float())
    err16_x0 = rel_l2(got.pred_original_sample.float(), want16.pred_original_sample.float())
    print(sched_name, "product-fp16 vs reference-in-fp16: prev", err16, "x0", err16_x0)
    assert err16 < 3e-3 and err16_x0 < 3e-3, (err16, err16_x0)
    assert abs(float(got.loss) - float(want16.loss)) < 3e-3 * float(want16.loss)
    if sched_name in ("dps", "mpgd"):
        err_prod = rel_l2(got.prev_sample.float(), want32.prev_sample)
        err_ref16 = rel_l2(want16.prev_sample.float(), want32.prev_sample)
        assert err_prod <= max(1.5 * err_ref16, 2e-3), (err_prod, err_ref16)


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("op_name", ["identity", "inpainting", "super_resolution", "super_resolution_x10",
                                     "phase_retrieval", "dereverberation"])
@pytest.mark.parametrize("space", ["mel_spectrogram", "wav_form"])
def test_16_bit_waveforms_are_consumed_directly(op_name, space, dt):
    """fused loss + VJP on a 16-bit waveform == the fp32 chain on the up-cast waveform, the gradient rounded once to the
    waveform dtype; also for a row-strided view and a preallocated strided gradient buffer (what the schedulers pass)."""
    B, L = 3, L1
    ops = dict(_ops(), identity=dm.IdentityOperator(16000),
               super_resolution_x10=dm.SuperResolutionOperator(sample_rate=16000, scale=10, noiser=_noiser()))
    op = ops[op_name]
    full = torch.zeros(B, L + 32, device=DEV, dtype=dt)
    full[:, :L] = stubs.synth_clips(B, L).to(DEV).to(dt)
    wav = full[:, :L]
    assert op.wave16_ok(wav)
    meas = op.forward(stubs.synth_clips(1, L, first=50).to(DEV))
    torch.manual_seed(5)  # dereverberation draws its impulse response inside the fused chain
    loss32, g32 = op.fused_loss_and_grad(wav.float(), meas, space)
    dfull = torch.full_like(full, 7.0)
    torch.manual_seed(5)
    loss16, g16 = op.fused_loss_and_grad(wav, meas, space, dwav=dfull[:, :L])
    assert g16.dtype == dt and g16.data_ptr() == dfull.data_ptr()
    assert torch.equal(loss16, loss32)
    assert torch.equal(g16, g32.to(dt))
    assert torch.all(dfull[:, L:] == 7.0)  # nothing written past L


def test_graphed_step_in_fp16_equals_eager_fp16():
    """the captured graph with 16-bit static latents (run.py:218 pipelines) replays the 16-bit kernel paths."""
    B, dt = 2, torch.float16
    vae, voc = stubs.StubVAE().to(DEV).to(dt), stubs.StubVocoder().to(DEV).to(dt)
    x, e = stubs.synth_latents(B, 25)
    x, e = x.to(DEV).to(dt), e.to(DEV).to(dt)
    op = _ops()["super_resolution"]
    sched = dm.get_scheduler("dps")(operator=op, **stubs.MUSICLDM_SCHED)
    sched.set_timesteps(500)
    meas = op.forward(stubs.synth_clips(1, L1, first=50).to(DEV))
    kw = dict(eta=0.0, measurement=meas, vae=vae, vocoder=voc, original_waveform_length=L1, ip_guidance_rate=5e-4,
              supervised_space="mel_spectrogram")
    graphed = dm.GraphedGuidedStep(sched, tuple(x.shape), dtype=dt, **kw)
    for t in (999, 501, 1):
        a = sched.step(e, t, x, **kw)
        b = graphed(e, t, x)
        assert b.prev_sample.dtype == dt
        assert torch.equal(a.prev_sample, b.prev_sample) and torch.equal(a.pred_original_sample, b.pred_original_sample)
        assert torch.equal(a.loss, b.loss) and torch.equal(a.loss_per_clip, b.loss_per_clip)


@pytest.mark.parametrize("dt", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("shape", [(16, 8, 250, 16), (3, 8, 25, 16), (2, 1, 1, 7), (5, 4, 33, 1031), (2, 8, 9000, 16)])
def test_batched_clip_noise_is_torch_randn_bit_for_bit(shape, dt):
    """dm_randn_clips == the reference's per-clip loop torch.randn((1, ...), generator=g_b) (torch_utils.py:31-76):
    same values, and the generators end in the same state (a second draw matches too)."""
    from diffmusic_b200 import ddim_base
    B = shape[0]
    ga = [torch.Generator(device=DEV).manual_seed(100 + i) for i in range(B)]
    gb = [torch.Generator(device=DEV).manual_seed(100 + i) for i in range(B)]
    ga[0].set_offset(40)  # a generator that has been used before
    gb[0].set_offset(40)
    per = (1,) + shape[1:]
    for _ in range(2):
        want = torch.cat([torch.randn(per, generator=g, device=DEV, dtype=dt) for g in ga])
        got = ddim_base.randn_clips_f32(shape, gb, torch.device(DEV), dt)
        assert got is not None and got.dtype == torch.float32
        assert torch.equal(got, want.float())
        assert [g.get_offset() for g in ga] == [g.get_offset() for g in gb]
    ddim_base.skip_randn(shape, gb, DEV, dt)
    torch.cat([torch.randn(per, generator=g, device=DEV, dtype=dt) for g in ga])
    assert [g.get_offset() for g in ga] == [g.get_offset() for g in gb]
    assert torch.equal(ddim_base.randn_tensor(shape, generator=gb, device=DEV, dtype=dt),
                       torch.cat([torch.randn(per, generator=g, device=DEV, dtype=dt) for g in ga]))
    # single generators and CPU generators keep going through torch
    assert ddim_base.randn_clips_f32(shape, gb[0], torch.device(DEV), dt) is None
    assert ddim_base.randn_clips_f32(shape, [torch.Generator().manual_seed(1) for _ in range(B)], torch.device(DEV), dt) is None


# ------------------------------------------------------------------------------------------------ batched-clip driver
@pytest.mark.parametrize("name,task,eta,rate,graph", [("dsg", "inpainting", 1.0, 0.08, False),
                                                      ("dps", "super_resolution", 0.0, 5e-4, False),
                                                      ("diffmusic", "inpainting", 1.0, 0.08, True)])
def test_batched_driver_matches_the_per_clip_pipeline_loop(name, task, eta, rate, graph):
    """BatchedGuidedSampler (4 clips in one batch, NaN guard evaluated once per trajectory, CUDA generators rewound to the
    offset right after the first NaN step) against the reference's structure: one clip per call, `torch.isnan(out.loss)`
    checked after every step, latents re-drawn from the clip's generator on a NaN (pipeline_musicldm.py:677-763).
    Clip 1 is poisoned at the third step of its first attempt."""
    from diffmusic_b200.ddim_base import randn_tensor
    from diffmusic_b200.driver import BatchedGuidedSampler
    torch.manual_seed(0)
    B, H, L, steps = 4, 25, 16000, 6
    nz = dm.get_noiser("gaussian", 0.0)
    op = (dm.MusicInpaintingOperator(1, 16000, "box", 0.2, 0.3, 0.3, 0.1, 1, noiser=nz) if task == "inpainting"
          else dm.SuperResolutionOperator(16000, scale=2, noiser=nz))
    sched = dm.get_scheduler(name)(operator=op, **stubs.MUSICLDM_SCHED)
    vae, voc = stubs.StubVAE().to(DEV), stubs.StubVocoder().to(DEV)
    net = torch.nn.Conv2d(8, 8, 3, padding=1).to(DEV)
    meas = torch.cat([op.forward(stubs.synth_clips(1, L, first=60 + j)) for j in range(B)]).to(DEV)  # run.py:286 per clip

    class Predict:
        def __init__(self):
            self.attempt = {}

        def __call__(self, x, t, clips):
            eps = net(x)
            for row, j in enumerate(clips):
                if int(t) == int(sched.timesteps[0]):
                    self.attempt[j] = self.attempt.get(j, 0) + 1
                if j == 1 and int(t) == int(sched.timesteps[2]) and self.attempt[j] == 1:
                    eps[row] = float("nan")
            return eps

    kw = dict(eta=eta, ip_guidance_rate=rate, supervised_space="mel_spectrogram")
    gens = [torch.Generator(device=DEV).manual_seed(500 + j) for j in range(B)]
    drv = BatchedGuidedSampler(sched, Predict(), vae, voc, num_inference_steps=steps, original_waveform_length=L,
                               latent_shape=(8, H, 16), graph=graph, **kw)
    out = drv(meas, gens, decode=True)
    assert out.restarts == [0, 1, 0, 0]
    assert torch.isfinite(out.latents).all() and torch.isfinite(out.loss_history).all()
    assert tuple(out.audios.shape) == (B, L) and out.audios.dtype == torch.float32

    # the reference's loop, one clip at a time
    for j in range(B):
        g = torch.Generator(device=DEV).manual_seed(500 + j)
        predict = Predict()
        sched.set_timesteps(steps, device=DEV)
        retry, restarts = 10, 0
        lat = randn_tensor((1, 8, H, 16), generator=g, device=DEV, dtype=torch.float32) * sched.init_noise_sigma
        while True:
            done = True
            for t in sched.timesteps:
                with torch.no_grad():
                    eps = predict(sched.scale_model_input(lat, t), t, [j])
                o = sched.step(eps, t, lat, generator=g, measurement=meas[j:j + 1], vae=vae, vocoder=voc,
                               original_waveform_length=L, **kw)
                if torch.isnan(o.loss) and retry >= 0:
                    retry -= 1
                    restarts += 1
                    lat = randn_tensor((1, 8, H, 16), generator=g, device=DEV, dtype=torch.float32) * sched.init_noise_sigma
                    done = False
                    break
                lat = o.prev_sample.detach()
            if done:
                break
        assert restarts == out.restarts[j]
        assert rel_l2(out.latents[j], lat[0]) < 1e-4, (j, rel_l2(out.latents[j], lat[0]))
        assert abs(float(out.loss[j]) - float(o.loss)) <= 1e-4 * abs(float(o.loss))
        assert gens[j].get_offset() == g.get_offset()  # same position in the clip's Philox stream, restarts included
