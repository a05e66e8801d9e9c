"""Batched-clip driver for the guided denoising loop.

The reference generates ONE clip per pipeline call: `get_dataloader(dataset, batch_size=1)` (run.py:249), a Python loop
over the files (run.py:264) and, inside `pipe(...)`, the denoising loop of pipelines/pipeline_musicldm.py:677-763 with a
host synchronisation on `torch.isnan(out.loss)` after every step (:742).  The trajectories of different clips are
independent (per-sample norms, per-clip generators), so this driver runs B of them as one batch through the
same scheduler / operator drop-ins:

  * per clip: its own measurement row, its own generator (the reference's list-of-generators semantics,
    torch_utils.py:31-76), its own restart counter;
  * the NaN guard keeps the reference's meaning -- a clip whose distance turns NaN is re-initialised from ITS generator and
    restarts ITS trajectory, at most `retry = 10` .. `0` = 11 times, after which the NaN run is accepted (:681,742-756) --
    but it is evaluated once per trajectory instead of once per step: a NaN is sticky (it poisons the clip's gradient,
    hence its latents, hence every later distance) and cannot leak into another clip, so the per-step distances are kept
    on the device and read back in one transfer at the end.  The clips that failed are re-run together as the next,
    smaller batch.  For CUDA generators the restart is bit-identical to what the step-by-step guard of a batch-1 run
    would have done: the generator is rewound to the Philox offset it had right after the first NaN step before the
    new latents are drawn.  (CPU generators cannot be rewound by offset; their restarts draw from the state at the end of
    the failed trajectory -- same distribution, different stream position.)
  * the step itself is `scheduler.step` (eager) or one CUDA-graph replay per step (`graph=True`, graph.py).

Networks stay what they are in the reference: arbitrary torch callables (`noise_predictor` = the UNet call with its
prompt conditioning and classifier-free guidance, `vae`, `vocoder`).  Multi-GPU: one process per GPU, clips sharded
`rank::world` (parallel.py), no collective.
"""
from __future__ import annotations

import inspect
from collections import OrderedDict
from dataclasses import dataclass, field
from typing import List, Optional

import torch

from .ddim_base import randn_tensor

MAX_RESTARTS = 11  # pipeline_musicldm.py:681 `retry = 10`, :742 restarts while `retry >= 0`


@dataclass
class BatchedSamplerOutput:
    latents: torch.Tensor                  # (B, C, H, W) final latents, in the clip order of the call
    loss: torch.Tensor                     # (B,) distance of the last step of the accepted trajectory
    loss_history: torch.Tensor             # (steps, B) per-step distances of the accepted trajectory
    restarts: List[int] = field(default_factory=list)  # NaN restarts per clip
    audios: Optional[torch.Tensor] = None  # (B, original_waveform_length) when decode=True


def _rewindable(gens):
    return isinstance(gens, (list, tuple)) and all(g.device.type == "cuda" for g in gens)


class BatchedGuidedSampler:
    """B guided trajectories per call instead of the reference's one clip per `pipe(...)`.

    scheduler        a diffmusic_b200 scheduler with its operator (`get_scheduler(name)(operator=op, **config)`)
    noise_predictor  callable `(latent_model_input, t) -> eps`; if it has a `clips` parameter it also receives the indices
                     (into the call's batch) of the clips in this sub-batch, for per-clip conditioning across restarts
    vae, vocoder     the differentiable decoders `scheduler.step` backpropagates through
    max_graphs       captured steps kept alive in graph mode (one per batch size and measurement tensor)
    step keywords    eta, ip_guidance_rate, supervised_space, original_waveform_length (+ eps for dsg / diffmusic through
                     `step_kwargs`): the keyword arguments of `scheduler.step` (pipeline_musicldm.py:726-739)
    """

    def __init__(self, scheduler, noise_predictor, vae, vocoder, *, num_inference_steps, original_waveform_length,
                 latent_shape=(8, 250, 16), eta=None, ip_guidance_rate=None, supervised_space="mel_spectrogram",
                 graph=False, dtype=torch.float32, max_restarts=MAX_RESTARTS, step_kwargs=None, max_graphs=2):
        self.scheduler = scheduler
        self.noise_predictor = noise_predictor
        self.vae, self.vocoder = vae, vocoder
        self.num_inference_steps = int(num_inference_steps)
        self.latent_shape = tuple(latent_shape)
        self.L = int(original_waveform_length)
        self.graph = bool(graph)
        self.dtype = dtype
        self.max_restarts = int(max_restarts)
        kw = dict(step_kwargs or {})
        if eta is not None:
            kw["eta"] = eta
        if ip_guidance_rate is not None:
            kw["ip_guidance_rate"] = ip_guidance_rate
        kw.update(supervised_space=supervised_space, original_waveform_length=self.L, vae=vae, vocoder=vocoder)
        self.step_kwargs = kw
        try:
            self._wants_clips = "clips" in inspect.signature(noise_predictor).parameters
        except (TypeError, ValueError):
            self._wants_clips = False
        self._graphs = OrderedDict()
        self.max_graphs = max(1, int(max_graphs))

    # ---- pieces of the reference pipeline -------------------------------------------------------------------------------
    def prepare_latents(self, generators, batch, device):
        """pipeline_musicldm.py:406-427: randn_tensor(shape, generator) * scheduler.init_noise_sigma."""
        if isinstance(generators, (list, tuple)) and len(generators) != batch:
            raise ValueError(f"You have passed a list of generators of length {len(generators)}, but requested an "
                             f"effective batch size of {batch}. Make sure the batch size matches the length of the "
                             f"generators.")
        z = randn_tensor((batch,) + self.latent_shape, generator=generators, device=device, dtype=self.dtype)
        return z * self.scheduler.init_noise_sigma

    def decode(self, latents):
        """pipeline_musicldm.py:768-781: 1/scaling_factor -> vae.decode -> vocoder -> [:, :L], fp32."""
        with torch.no_grad():
            mel = self.vae.decode(1 / self.vae.config.scaling_factor * latents).sample
            if mel.dim() == 4:
                mel = mel.squeeze(1)
            return self.vocoder(mel)[:, :self.L].float()

    def _predict(self, x, t, ids):
        with torch.no_grad():
            if self._wants_clips:
                return self.noise_predictor(x, t, clips=ids)
            return self.noise_predictor(x, t)

    def _stepper(self, batch, measurement, device):
        if not self.graph:
            kw = dict(self.step_kwargs, measurement=measurement)
            return lambda eps, t, x, gens: self.scheduler.step(eps, t, x, generator=gens, **kw)
        from .graph import GraphedGuidedStep
        key = (batch, measurement.data_ptr(), tuple(measurement.shape))
        g = self._graphs.get(key)
        if g is None:
            # a captured step holds its measurement, static buffers and the networks' activations: keep only the most
            # recent ones (normally the full batch and a restart sub-batch).  Every trajectory ends with a host read of
            # its distances, so no replay of an evicted graph is still in flight here.
            while len(self._graphs) >= self.max_graphs:
                self._graphs.popitem(last=False)
            g = self._graphs[key] = GraphedGuidedStep(self.scheduler, (batch,) + self.latent_shape, dtype=self.dtype,
                                                      device=device, measurement=measurement, **self.step_kwargs)
        else:
            self._graphs.move_to_end(key)
        return lambda eps, t, x, gens: g(eps, t, x, generator=gens)

    # ---- one trajectory of a (sub-)batch ---------------------------------------------------------------------------------
    def _trajectory(self, ids, measurement, gens, latents):
        sched = self.scheduler
        timesteps = sched.timesteps
        dev = latents.device
        step = self._stepper(len(ids), measurement, dev)
        hist = torch.zeros((len(timesteps), len(ids)), device=dev, dtype=torch.float32)
        rewind = _rewindable(gens)
        start = per_step = None
        # one host read of the schedule per trajectory: the step needs Python ints (coefficient lookup), and reading
        # them one by one from the device tensor would synchronise the host with the stream at every step
        t_ints = timesteps.tolist()
        for i, t_int in enumerate(t_ints):
            t = timesteps[i]  # the noise predictor still gets the tensor element the reference pipelines pass
            eps = self._predict(sched.scale_model_input(latents, t), t, ids)
            if i == 0 and rewind:
                start = [g.get_offset() for g in gens]
            out = step(eps, t_int, latents, gens)
            if i == 0 and rewind:  # every step consumes the same amount of every clip's stream
                per_step = [g.get_offset() - s for g, s in zip(gens, start)]
            per_clip = getattr(out, "loss_per_clip", None)
            if per_clip is not None:
                hist[i].copy_(per_clip.reshape(-1))
            elif out.loss is not None and out.loss.numel() == 1 and out.loss.is_floating_point():
                hist[i].fill_(0.0).add_(out.loss.reshape(()).to(hist.dtype))  # whole-batch scalar (foreign scheduler)
            latents = out.prev_sample.detach()
        return latents, hist, start, per_step

    # ---- the driver ----------------------------------------------------------------------------------------------------
    def __call__(self, measurement, generators=None, *, batch=None, device=None, latents=None, decode=False,
                 clip_ids=None):
        """measurement: (1, ...) shared by the batch or (B, ...) one row per clip (what `operator.forward(clip)` returned,
        run.py:286,312); generators: list of B generators (or one generator / None, then restarts cannot isolate clips);
        latents: optional (B, ...) initial latents (first attempt only, like `pipe(latents=...)`); clip_ids: the labels the
        noise predictor receives as `clips` (default 0..B-1; a sharded run passes the global clip indices)."""
        sched = self.scheduler
        if isinstance(generators, (list, tuple)):
            B = len(generators)
        elif latents is not None:
            B = latents.shape[0]
        elif batch is not None:
            B = int(batch)
        else:
            B = measurement.shape[0]
        if measurement.shape[0] not in (1, B):
            raise ValueError(f"measurement holds {measurement.shape[0]} rows for a batch of {B} clips")
        dev = torch.device(device) if device is not None else measurement.device
        measurement = measurement.to(dev)
        sched.set_timesteps(self.num_inference_steps, device=dev)
        steps = len(sched.timesteps)
        per_clip_gens = isinstance(generators, (list, tuple))
        labels = list(range(B)) if clip_ids is None else list(clip_ids)
        if len(labels) != B:
            raise ValueError(f"clip_ids names {len(labels)} clips for a batch of {B}")

        final = [None] * B
        final_hist = [None] * B
        restarts = [0] * B
        queue = list(range(B))
        first = True
        while queue:
            ids = queue
            gens = [generators[j] for j in ids] if per_clip_gens else generators
            meas = measurement if measurement.shape[0] == 1 or len(ids) == B else measurement[ids].contiguous()
            if first and latents is not None:
                x = latents.to(dev) * sched.init_noise_sigma
            else:
                x = self.prepare_latents(gens, len(ids), dev)
            first = False
            x, hist, start, per_step = self._trajectory([labels[j] for j in ids], meas, gens, x)
            bad = torch.isnan(hist)
            first_bad = torch.where(bad.any(0), bad.float().argmax(0), torch.full_like(bad[0], -1, dtype=torch.long))
            first_bad = first_bad.tolist()  # the one host synchronisation of the trajectory
            queue = []
            for col, j in enumerate(ids):
                if first_bad[col] >= 0 and restarts[j] < self.max_restarts:
                    restarts[j] += 1
                    if start is not None:  # rewind to the state right after the first NaN step (batch-1 semantics)
                        gens[col].set_offset(start[col] + (first_bad[col] + 1) * per_step[col])
                    queue.append(j)
                else:
                    final[j] = x[col]
                    final_hist[j] = hist[:, col]
        out_latents = torch.stack(final)
        out_hist = torch.stack(final_hist, dim=1) if steps else torch.zeros((0, B), device=dev)
        result = BatchedSamplerOutput(latents=out_latents, loss=out_hist[-1] if steps else torch.zeros(B, device=dev),
                                      loss_history=out_hist, restarts=restarts)
        if decode:
            result.audios = self.decode(out_latents)
        return result
