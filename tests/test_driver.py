"""Host logic of the batched-clip driver (diffmusic_b200/driver.py) with a scheduler stand-in that
runs on the CPU: queueing, per-clip isolation of the NaN restart guard (pipeline_musicldm.py:681,742-756), measurement
rows, restart cap.  The real schedulers / kernels are exercised by tests/test_gpu_parity.py."""
import pytest
import torch

from diffmusic_b200.ddim_base import randn_tensor
from diffmusic_b200.driver import MAX_RESTARTS, BatchedGuidedSampler
from diffmusic_b200.schedulers import InverseProblemSchedulerOutput

SHAPE = (2, 5, 4)


class ToyScheduler:
    """same call surface as the drop-in schedulers; elementwise arithmetic, so a batch equals its clips bit for bit"""
    init_noise_sigma = 1.5

    def __init__(self):
        self.seen_measurements = []

    def set_timesteps(self, n, device=None):
        self.timesteps = torch.arange(n - 1, -1, -1)

    def scale_model_input(self, x, t=None):
        return x

    def step(self, eps, t, x, generator=None, measurement=None, eta=0.0, **kw):
        z = randn_tensor(x.shape, generator=generator, device=x.device, dtype=x.dtype) if eta > 0 else torch.zeros_like(x)
        prev = 0.9 * x - 0.1 * eps + 0.05 * eta * z
        self.seen_measurements.append(measurement.clone())
        loss = (prev.reshape(x.shape[0], -1).mean(1, keepdim=True) - measurement.reshape(measurement.shape[0], -1)
                ).norm(dim=1).expand(x.shape[0]).clone()
        return InverseProblemSchedulerOutput(prev_sample=prev, pred_original_sample=prev, loss=loss.sum(),
                                             loss_per_clip=loss)


class Predictor:
    """eps = tanh(x); returns NaN for the clips in `poison` at step `at`, on their first `times` attempts"""

    def __init__(self, poison=(), at=3, times=1):
        self.poison, self.at, self.times = set(poison), at, times
        self.attempt = {}

    def __call__(self, x, t, clips):
        eps = torch.tanh(x)
        for row, j in enumerate(clips):
            if int(t) == 7:  # first timestep of the 8-step schedule: a new attempt of clip j starts
                self.attempt[j] = self.attempt.get(j, 0) + 1
            if j in self.poison and int(t) == self.at and self.attempt[j] <= self.times:
                eps[row] = float("nan")
        return eps


def sampler(pred, **kw):
    return BatchedGuidedSampler(ToyScheduler(), pred, None, None, num_inference_steps=8, original_waveform_length=0,
                                latent_shape=SHAPE, **kw)


def gens(ids):
    return [torch.Generator().manual_seed(40 + j) for j in ids]


def test_batch_equals_its_clips_and_restarts_stay_per_clip():
    meas = torch.arange(4, dtype=torch.float32).reshape(4, 1) * 0.1
    clean = sampler(Predictor(), eta=1.0)(meas, gens(range(4)))
    assert clean.restarts == [0, 0, 0, 0] and clean.latents.shape == (4,) + SHAPE
    assert clean.loss_history.shape == (8, 4) and torch.equal(clean.loss, clean.loss_history[-1])
    for j in range(4):  # one clip per call, as the reference's loader loop does
        one = sampler(Predictor(), eta=1.0)(meas[j:j + 1], gens([j]))
        assert torch.equal(one.latents[0], clean.latents[j])
        assert torch.equal(one.loss_history[:, 0], clean.loss_history[:, j])
    # clip 2 turns NaN at t = 3 on its first attempt: only clip 2 restarts, the others keep their trajectories
    s = sampler(Predictor(poison=[2]), eta=1.0)
    out = s(meas, gens(range(4)))
    assert out.restarts == [0, 0, 1, 0]
    assert torch.isfinite(out.latents).all() and torch.isfinite(out.loss_history).all()
    for j in (0, 1, 3):
        assert torch.equal(out.latents[j], clean.latents[j])
    assert not torch.equal(out.latents[2], clean.latents[2])  # new latents from the clip's own generator
    # the retried sub-batch saw clip 2's own measurement row
    assert torch.equal(s.scheduler.seen_measurements[-1], meas[2:3])
    # and equals the same clip driven alone through the same failure
    alone = sampler(Predictor(poison=[0]), eta=1.0)(meas[2:3], gens([2]))
    assert alone.restarts == [1] and torch.equal(alone.latents[0], out.latents[2])


def test_restart_cap_accepts_the_nan_run_like_the_reference():
    meas = torch.zeros(1, 1)
    out = sampler(Predictor(poison=[1], times=10 ** 6), eta=0.0)(meas, gens(range(3)))
    assert out.restarts == [0, MAX_RESTARTS, 0] and MAX_RESTARTS == 11
    assert torch.isnan(out.latents[1]).all() and torch.isfinite(out.latents[[0, 2]]).all()
    few = sampler(Predictor(poison=[1], times=2), eta=0.0, max_restarts=5)(meas, gens(range(3)))
    assert few.restarts == [0, 2, 0] and torch.isfinite(few.latents).all()


def test_arguments():
    s = sampler(Predictor())
    with pytest.raises(ValueError):
        s(torch.zeros(3, 1), gens(range(4)))
    with pytest.raises(ValueError):
        s.prepare_latents(gens(range(2)), 3, torch.device("cpu"))
    given = torch.ones((2,) + SHAPE)
    out = s(torch.zeros(1, 1), gens(range(2)), latents=given)
    ref = sampler(Predictor())(torch.zeros(1, 1), gens(range(2)))
    assert not torch.equal(out.latents, ref.latents)
    # a single shared generator draws the whole batch at once (torch_utils.py:31-76)
    g = torch.Generator().manual_seed(5)
    shared = sampler(lambda x, t: torch.tanh(x))(torch.zeros(1, 1), g, batch=3)
    assert shared.latents.shape == (3,) + SHAPE


def test_graph_mode_keeps_a_bounded_number_of_captured_steps(monkeypatch):
    """graph=True builds one captured step per (batch size, measurement tensor); the cache is an LRU of `max_graphs`."""
    import diffmusic_b200.graph as graph_mod
    built = []

    class FakeGraphedStep:
        def __init__(self, scheduler, shape, *, dtype, device, measurement, **kw):
            self.sched, self.kw = scheduler, dict(kw, measurement=measurement)
            built.append((shape[0], measurement.data_ptr()))

        def __call__(self, eps, t, x, generator=None):
            kw = {k: v for k, v in self.kw.items() if k not in ("vae", "vocoder")}
            return self.sched.step(eps, t, x, generator=generator, **kw)

    monkeypatch.setattr(graph_mod, "GraphedGuidedStep", FakeGraphedStep)
    s = sampler(Predictor(poison=[1]), eta=1.0, graph=True, max_graphs=2)
    meas = [torch.full((3, 1), float(i)) for i in range(3)]
    eager = sampler(Predictor(poison=[1]), eta=1.0)(meas[0], gens(range(3)))
    out = s(meas[0], gens(range(3)))
    assert out.restarts == [0, 1, 0] and torch.equal(out.latents, eager.latents)
    assert [b for b, _ in built] == [3, 1] and len(s._graphs) == 2  # full batch + the restart sub-batch
    s.noise_predictor = Predictor()
    s(meas[0], gens(range(3)))
    assert len(built) == 2  # same measurement tensor, same batch: the captured step is reused
    s(meas[1], gens(range(3)))
    s(meas[2], gens(range(3)))
    assert len(built) == 4 and len(s._graphs) == 2
    assert [k[1] for k in s._graphs] == [meas[1].data_ptr(), meas[2].data_ptr()]
