"""Pin the CPU oracle (oracle/) to fixtures produced by the reference's own Python (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import operators as oo
from oracle import steps as osteps
from tests import stubs
from tests.conftest import rel_l2

L1, LP = 16000, 4000
TOL = 1e-6  # same library calls as the reference on the same CPU -> agreement to fp32 rounding


def _loss_grad(op, wav, meas, space):
    w = wav.clone().requires_grad_(True)
    pred = op.forward(w)
    diff = (meas - pred) if space == "wav_form" else (op.transform(meas) - op.transform(pred))
    loss = torch.linalg.norm(diff)
    return loss.detach(), torch.autograd.grad(loss, w)[0]


def test_masks_bit_exact(golden_ops):
    kw = dict(audio_length_in_s=10, sample_rate=16000, mask_percentage=0.3, interval_s=1, mask_duration_s=0.1)
    cases = {
        "mask_box_10s": dict(mask_type="box", start_inpainting_s=2, end_inpainting_s=3),
        "mask_boxfrac_10s": dict(mask_type="box", start_inpainting_s=1.37, end_inpainting_s=2.913),
        "mask_periodic_10s": dict(mask_type="periodic"),
    }
    for key, c in cases.items():
        m = oo.inpaint_mask(**kw, **c)
        assert np.array_equal(np.packbits(m[0].numpy().astype(np.uint8)), golden_ops[key]), key
    torch.manual_seed(7)
    m = oo.inpaint_mask(mask_type="random", **kw)
    assert np.array_equal(np.packbits(m[0].numpy().astype(np.uint8)), golden_ops["mask_random_seed7_10s"])
    box = np.unpackbits(golden_ops["mask_box_10s"])[:160000]
    assert box.sum() == 160000 - 16000 and not box[32000:48000].any()  # SURVEY 8c known answer


def test_transforms_and_forwards(golden_ops):
    wav = stubs.synth_clips(2, L1)
    mask = oo.inpaint_mask(1, 16000, "box", 0.25, 0.5)
    assert rel_l2(oo.mel_db(wav), golden_ops["identity_transform"]) < TOL
    assert rel_l2(oo.a_inpaint(wav, mask), golden_ops["inpaint_forward"]) == 0.0
    assert rel_l2(oo.mel_db(oo.a_inpaint(wav, mask), clamp=False), golden_ops["inpaint_transform"]) < TOL
    for s in (2, 10):
        y = oo.a_superres(wav, 16000, s)
        assert rel_l2(y, golden_ops[f"superres_forward_s{s}"]) < TOL
        assert rel_l2(oo.mel_db(y), golden_ops[f"superres_transform_s{s}"]) < TOL
    mag = oo.a_phase(wav[:, :LP])
    assert rel_l2(mag, golden_ops["phase_forward"]) < TOL
    assert rel_l2(oo.phase_mel(mag), golden_ops["phase_transform"]) < TOL
    for K, decay in ((800, 0.85), (5000, 0.99), (801, 0.9)):
        torch.manual_seed(100 + K)
        ir = oo.draw_impulse_response(K, decay)
        assert np.array_equal(ir.numpy(), golden_ops[f"dereverb_ir_K{K}"])
        assert rel_l2(oo.a_dereverb(wav, ir), golden_ops[f"dereverb_forward_K{K}"]) < 1e-5
    torch.manual_seed(5)
    assert np.array_equal(oo.gaussian_noise(wav[:1, :256], 0.05).numpy(), golden_ops["gaussian_noise_s0.05_seed5"])


@pytest.mark.parametrize("space", ["mel_spectrogram", "wav_form"])
def test_loss_and_grad(golden_ops, space):
    wav = stubs.synth_clips(1, L1)
    ref = stubs.synth_clips(1, L1, first=50)
    mask = oo.inpaint_mask(1, 16000, "box", 0.25, 0.5)
    ops = {"inpaint": oo.OracleOperator("inpainting", mask=mask),
           "superres_s2": oo.OracleOperator("super_resolution", scale=2),
           "superres_s10": oo.OracleOperator("super_resolution", scale=10)}
    for K in (800, 5000, 801):
        ops[f"dereverb_K{K}"] = oo.OracleOperator("dereverberation",
                                                  fixed_ir=torch.from_numpy(golden_ops[f"dereverb_ir_K{K}"]))
    for name, op in ops.items():
        l, g = _loss_grad(op, wav, op.forward(ref), space)
        assert abs(float(l) - float(golden_ops[f"{name}_{space}_loss"])) <= 2e-5 * abs(float(l)), name
        assert rel_l2(g, golden_ops[f"{name}_{space}_grad"]) < 2e-4, name
    ph = oo.OracleOperator("phase_retrieval")
    l, g = _loss_grad(ph, wav[:, :LP], ph.forward(ref[:, :LP]), space)
    assert abs(float(l) - float(golden_ops[f"phase_{space}_loss"])) <= 2e-5 * abs(float(l))
    assert rel_l2(g, golden_ops[f"phase_{space}_grad"]) < 2e-4


def _step_cases(golden_steps):
    seen = []
    for k in golden_steps.files:
        if k.endswith("|prev"):
            seen.append(k[:-5])
    return seen


def test_scheduler_constants(golden_steps):
    base = osteps.make_base(**stubs.MUSICLDM_SCHED)
    base.set_timesteps(500)
    assert np.array_equal(base.timesteps.numpy(), golden_steps["timesteps_500"])
    assert base.timesteps[0] == 999 and base.timesteps[-1] == 1
    assert np.array_equal(base.alphas_cumprod.numpy(), golden_steps["alphas_cumprod"])
    assert abs(float(base.final_alpha_cumprod) - 0.99849999) < 1e-7      # SURVEY 8c known answers
    assert abs(float(base.alphas_cumprod[999]) - 1.4230386e-4) < 1e-10


def test_steps_match_reference(golden_steps):
    vae, voc = stubs.StubVAE(), stubs.StubVocoder()
    ref_wav = stubs.synth_clips(1, L1, first=50)
    x, e = stubs.synth_latents(1, 25)
    mask = oo.inpaint_mask(1, 16000, "box", 0.25, 0.5)
    ops = {"inpainting": oo.OracleOperator("inpainting", mask=mask),
           "super_resolution": oo.OracleOperator("super_resolution", scale=2),
           "phase_retrieval": oo.OracleOperator("phase_retrieval"),
           "dereverberation": oo.OracleOperator("dereverberation", ir_length=800, decay_factor=0.85),
           "identity": oo.OracleOperator("identity")}
    base = osteps.make_base(**stubs.MUSICLDM_SCHED)
    base.set_timesteps(500)
    cases = _step_cases(golden_steps)
    assert len(cases) == 33
    for key in cases:
        sched, op_name, space, eta, t = key.split("|")
        eta, t = float(eta[3:]), int(t[1:])
        op = ops[op_name]
        torch.manual_seed(321)
        meas = op.forward(ref_wav)
        rate = {"ddim": None, "dps": 5e-4, "mpgd": 0.005, "dsg": 0.08, "diffmusic": 0.08}[sched]
        gen = torch.Generator().manual_seed(3000)
        torch.manual_seed(654 + t)
        o = osteps.reference_step(sched, base, op, e, t, x, eta=eta, ip_guidance_rate=rate, generator=gen,
                                  measurement=meas, vae=vae, vocoder=voc, original_waveform_length=L1,
                                  supervised_space=space)
        assert rel_l2(o.prev_sample, golden_steps[key + "|prev"]) < 5e-6, key
        assert rel_l2(o.pred_original_sample, golden_steps[key + "|x0"]) < 5e-6, key
        gl = float(golden_steps[key + "|loss"].ravel()[0])
        assert abs(float(o.loss.float().ravel()[0]) - gl) <= 1e-5 * max(1.0, abs(gl)), key


def test_clip_sample_steps_match_reference(golden_steps_clip):
    """clip_sample=True: the oracle differentiates through the clamp of the base step like the reference
    (scheduling_dps.py:165-175); 4 schedulers x t in {999, 501, 1}, 18-99 % of the latent clipped."""
    vae, voc = stubs.StubVAE(), stubs.StubVocoder()
    ref_wav = stubs.synth_clips(1, L1, first=50)
    x, e = stubs.synth_latents(1, 25)
    op = oo.OracleOperator("inpainting", mask=oo.inpaint_mask(1, 16000, "box", 0.25, 0.5))
    base = osteps.make_base(**dict(stubs.MUSICLDM_SCHED, clip_sample=True))
    base.set_timesteps(500)
    cases = _step_cases(golden_steps_clip)
    assert len(cases) == 12
    meas = op.forward(ref_wav)
    for key in cases:
        sched, _, space, eta, t = key.split("|")
        eta, t = float(eta[3:]), int(t[1:])
        rate = {"dps": 5e-4, "mpgd": 0.005, "dsg": 0.08, "diffmusic": 0.08}[sched]
        o = osteps.reference_step(sched, base, op, e, t, x, eta=eta, ip_guidance_rate=rate,
                                  generator=torch.Generator().manual_seed(3000), measurement=meas, vae=vae, vocoder=voc,
                                  original_waveform_length=L1, supervised_space=space)
        assert rel_l2(o.prev_sample, golden_steps_clip[key + "|prev"]) < 5e-6, key
        assert rel_l2(o.pred_original_sample, golden_steps_clip[key + "|x0"]) < 5e-6, key
        gl = float(golden_steps_clip[key + "|loss"].ravel()[0])
        assert abs(float(o.loss.float().ravel()[0]) - gl) <= 1e-5 * max(1.0, abs(gl)), key


# ------------------------------------------------------------------------------------------------ section 8(f) oracles
def test_frechet_oracle_closed_forms():
    """oracle.fad.calc_frechet_distance (restated fadtk/fad.py:50-119) against closed forms: identical Gaussians -> 0;
    covariances sharing an eigenbasis -> |dmu|^2 + sum_i (sqrt(a_i) - sqrt(b_i))^2."""
    from oracle import fad as ofad
    rng = np.random.default_rng(0)
    d = 24
    q, _ = np.linalg.qr(rng.standard_normal((d, d)))
    a, b = rng.uniform(0.1, 3.0, d), rng.uniform(0.1, 3.0, d)
    c1, c2 = (q * a) @ q.T, (q * b) @ q.T
    mu1, mu2 = rng.standard_normal(d), rng.standard_normal(d)
    want = np.sum((mu1 - mu2) ** 2) + np.sum((np.sqrt(a) - np.sqrt(b)) ** 2)
    got = float(np.real(ofad.calc_frechet_distance(mu1, c1, mu2, c2)))
    assert abs(got - want) < 1e-9 * want
    assert abs(float(np.real(ofad.calc_frechet_distance(mu1, c1, mu1, c1)))) < 1e-9
    np.random.seed(3)
    emb = rng.standard_normal((900, d)) * np.sqrt(a)
    score, slope, r2, pts = ofad.score_inf(mu1 * 0, c1, emb, steps=5, min_n=100)
    assert [p[0] for p in pts] == [100, 300, 500, 700, 900] and np.isfinite([score, slope, r2]).all()
    assert slope > 0  # the finite-sample bias of FAD shrinks like 1/n


def test_fad_oracle_matches_reference_functions():
    """oracle.fad against tests/golden/fad.npz = the outputs of the reference's own calc_embd_statistics,
    calc_frechet_distance, score_inf (fadtk/fad.py:41-47, 50-119, 303-350) and calculate_embd_statistics_online
    (fadtk/utils.py:13-46), cut out of the reference with `ast` and run unmodified (tests/golden/make_fad_golden.py)."""
    import os
    from oracle import fad as ofad
    from tests.conftest import GOLDEN
    z = np.load(os.path.join(GOLDEN, "fad.npz"))
    for name, (n1, n2, d, parts) in stubs.FAD_CASES.items():
        a, b = stubs.fad_embeddings(name)
        mu1, c1 = ofad.calc_embd_statistics(a)
        mu2, c2 = ofad.calc_embd_statistics(b)
        assert mu1.dtype == np.float16 and np.array_equal(mu1, z[name + "_mu1"])   # the fp16 mean of SURVEY.md D.11
        assert np.array_equal(mu2, z[name + "_mu2"])
        assert np.array_equal(c1, z[name + "_cov1"]) and np.array_equal(c2, z[name + "_cov2"])
        fd = float(np.real(ofad.calc_frechet_distance(mu1, c1, mu2, c2)))
        assert abs(fd - float(z[name + "_fd"])) <= 1e-9 * float(z[name + "_fd"])
        omu, ocov = ofad.embd_statistics_online(np.array_split(a, parts))
        assert np.array_equal(omu, z[name + "_online_mu"]) and np.array_equal(ocov, z[name + "_online_cov"])
    a, b = stubs.fad_embeddings("d128")
    mu_b, cov_b = ofad.calc_embd_statistics(b)
    np.random.seed(1234)
    score, slope, r2, points = ofad.score_inf(mu_b, cov_b, a, steps=6, min_n=200)
    want = z["inf_points"]
    assert np.array_equal(np.array(points)[:, 0], want[:, 0])
    assert np.allclose(np.real(np.array(points)[:, 1]), want[:, 1], rtol=1e-9, atol=0)
    assert abs(score - float(z["inf_score"])) <= 1e-8 * abs(float(z["inf_score"]))
    assert abs(r2 - float(z["inf_r2"])) <= 1e-8


def test_metric_oracles_closed_forms():
    """oracle.metrics: LSD of a clip with itself is 0, scaling a clip by c shifts every log-magnitude by log10(c);
    MSE of constant offsets; nan_to_num sanitising as in lsd.py:23 / mse.py:14-15."""
    from oracle import metrics as om
    x = stubs.synth_clips(2, 8000).numpy()
    assert om.lsd_score(x, x, 1024, 512) == 0.0
    got = om.lsd_score(x, 10.0 * x, 1024, 512, eps=0.0)
    assert abs(got - 1.0) < 1e-4
    y = x.copy()
    y[0, 5] = np.nan
    y[1, 9] = np.inf
    z = x.copy()
    z[0, 5], z[1, 9] = 0.0, 1.0
    assert om.lsd_score(x, y, 1024, 512) == om.lsd_score(x, z, 1024, 512)
    assert abs(om.mse_score(x, x + 0.5) - 0.25) < 1e-6
    assert abs(om.mse_score(x, x + 0.5, "sum") - 0.5) < 1e-6
    assert om.mse_score(x, y) == om.mse_score(x, z)


@pytest.mark.parametrize("name", ["b1_t26", "b3_t41_own_phase", "b2_t9"])
def test_istft_oracle_matches_reference(name):
    """oracle/istft.py against the reference's own mel_spectrogram_to_waveform_with_phase (pipeline_musicldm.py:263-301,
    run in the build container by tests/golden/make_istft_golden.py): InverseMelScale as the minimum-norm matrix, the
    rectangular-window istft, clip / zero-pad to original_waveform_length."""
    import os
    from oracle import istft as oi
    from tests.conftest import GOLDEN
    want = np.load(os.path.join(GOLDEN, "istft.npz"))[name]
    mel, phase = stubs.istft_inputs(name)
    got = oi.mel_spectrogram_to_waveform_with_phase(mel.numpy(), phase.numpy(),
                                                    original_waveform_length=stubs.ISTFT_CASES[name][4])
    assert got.shape == want.shape
    assert rel_l2(got, want) < 2e-6  # the reference runs lstsq + irfft in fp32
    n = 160 * (stubs.ISTFT_CASES[name][1] - 1)
    if want.shape[1] > n:
        assert not want[:, n:].any() and not got[:, n:].any()


def test_inverse_mel_scale_is_the_minimum_norm_matrix():
    """live torchaudio InverseMelScale (lstsq, driver gels) == relu(fb (fb^T fb)^-1 mel), and the host table the kernel
    reads (diffmusic_b200.tables.inverse_mel_matrix) is that matrix."""
    import torchaudio
    from diffmusic_b200 import tables
    from oracle import istft as oi
    g = torch.Generator().manual_seed(5)
    mel = torch.rand(2, 64, 33, generator=g) * 8.0 - 2.0
    live = torchaudio.transforms.InverseMelScale(n_stft=513, n_mels=64, sample_rate=16000)(mel)
    assert rel_l2(oi.inverse_mel_scale(mel.numpy()), live) < 2e-6
    w_t = tables.inverse_mel_matrix(16000)
    assert w_t.shape == (64, 513) and w_t.dtype == torch.float32
    assert rel_l2(torch.relu(torch.einsum("mk,bmt->bkt", w_t, mel)), live) < 2e-6
    assert not w_t[:, 0].any() and not w_t[:, 512].any()  # the filterbank has no weight on DC / Nyquist


def _spectrogram_error(mag, phase, want_mag, want_phase):
    """relative L2 error of mag * exp(i phase) (the phase alone is ill-conditioned where the magnitude vanishes and jumps
    by 2 pi at the branch cut), and the largest wrapped phase error over the bins that carry energy"""
    mag, phase, want_mag, want_phase = (np.asarray(a, np.float64) for a in (mag, phase, want_mag, want_phase))
    z, w = mag * np.exp(1j * phase), want_mag * np.exp(1j * want_phase)
    rel = np.linalg.norm(z - w) / np.linalg.norm(w)
    loud = want_mag > 1e-3 * want_mag.max()
    d = np.angle(np.exp(1j * (phase - want_phase)))
    return rel, np.abs(d[loud]).max()


@pytest.mark.parametrize("name", sorted(stubs.SPECTROGRAM_CASES))
def test_waveform_to_spectrogram_oracle_matches_reference(name):
    """oracle/istft.py waveform_to_spectrogram against the reference function's own output (diffmusic/utils.py:11-20 run
    by tests/golden/make_istft_golden.py)."""
    import os
    from oracle import istft as oi
    from tests.conftest import GOLDEN
    z = np.load(os.path.join(GOLDEN, "istft.npz"))
    wav = stubs.synth_clips(2, stubs.SPECTROGRAM_CASES[name], first=70).numpy()
    mag, phase = oi.waveform_to_spectrogram(wav)
    assert mag.shape == z[name + "_mag"].shape
    rel, dphi = _spectrogram_error(mag, phase, z[name + "_mag"], z[name + "_phase"])
    assert rel < 2e-6 and dphi < 2e-3
    assert rel_l2(mag, z[name + "_mag"]) < 2e-6


@pytest.mark.parametrize("hop,T", [(160, 12), (512, 37), (1024, 5), (160, 2)])
def test_istft_oracle_vs_live_libraries(hop, T):
    """The GPU parity tests use oracle/istft.py at hops the reference fixture does not cover: hold it to the live
    torchaudio InverseMelScale + torch.istft (the calls the reference makes) there as well."""
    import torchaudio
    from oracle import istft as oi
    g = torch.Generator().manual_seed(100 + hop + T)
    mel = torch.rand(2, 1, T, 64, generator=g) * 5.0 - 0.5
    phase = (torch.rand(2, 513, T, generator=g) * 2.0 - 1.0) * np.pi
    lin = torchaudio.transforms.InverseMelScale(n_stft=513, n_mels=64, sample_rate=16000)(mel.squeeze(1).permute(0, 2, 1))
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = torch.istft(lin * torch.exp(1j * phase), n_fft=1024, hop_length=hop, win_length=1024)
    got = oi.mel_spectrogram_to_waveform_with_phase(mel.numpy(), phase.numpy(), hop_length=hop)
    assert got.shape == tuple(want.shape) == (2, hop * (T - 1))
    assert rel_l2(got, want) < 3e-6


def test_waveform_to_spectrogram_oracle_vs_live_torch():
    from oracle import istft as oi
    wav = stubs.synth_clips(2, 20000, first=90)
    spec = torch.stft(wav, 1024, hop_length=512, win_length=1024, return_complex=True)
    mag, phase = oi.waveform_to_spectrogram(wav.numpy(), hop_length=512)
    rel, dphi = _spectrogram_error(mag, phase, spec.abs().numpy(), spec.angle().numpy())
    assert mag.shape == tuple(spec.shape) and rel < 2e-6 and dphi < 2e-3
