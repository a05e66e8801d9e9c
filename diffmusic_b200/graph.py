"""CUDA-graph capture of a whole guided step.

One guided step moves ~2-3 MB per clip and is launch-bound (SURVEY.md 7.3 "tiny problem sizes", 8(f) rank 4): ~40
kernel launches (ours + the torch networks and their autograd) cost more host time than device time.
`GraphedGuidedStep` captures `scheduler.step(...)` once -- x0 kernel, vae.decode, vocoder, fused operator/loss/VJP
kernels, torch autograd back through the networks, fused update kernel -- and replays it for every timestep:

  * latents go through static input buffers; the per-timestep scalars are 24 bytes in a device coefficient vector the
    scheduler kernels read (`coef` argument of dm_sched_*), so nothing timestep-dependent is baked into the graph;
  * all randomness stays in torch and outside the graph: the step noise is drawn before the replay with exactly the
    draws, order and generator semantics of the eager step (`draw_step_noise`), the dereverberation impulse response is
    drawn on the host as the reference does (operator.py:238-242) and copied into a static buffer;
  * nothing inside the step synchronises with the host (the slerp branch is resolved on the device).
"""
from __future__ import annotations

import torch

from .schedulers import InverseProblemSchedulerOutput


class GraphedGuidedStep:
    """Callable with the call shape of `scheduler.step(model_output, timestep, sample, generator=...)`.

    step_kwargs are the fixed keyword arguments of the step: eta, measurement, vae, vocoder,
    original_waveform_length, ip_guidance_rate, supervised_space, eps.
    Returned tensors are fresh copies unless clone_outputs=False (then they are valid until the next call).
    """

    def __init__(self, scheduler, sample_shape, *, dtype=torch.float32, device=None, clone_outputs=True,
                 warmup_timestep=None, **step_kwargs):
        self.sched = scheduler
        self.kw = dict(step_kwargs)
        self.eta = float(self.kw.get("eta", _default_eta(scheduler)))
        self.kw["eta"] = self.eta
        self.clone = clone_outputs
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.x = torch.zeros(sample_shape, device=dev, dtype=dtype)
        self.e = torch.zeros(sample_shape, device=dev, dtype=dtype)
        self.B = sample_shape[0]
        self.n_clip = self.x.numel() // self.B
        needs_noise = scheduler.noise_mode == "always" or (self.eta > 0 and type(scheduler).__name__ != "DDIMScheduler")
        self.z = torch.zeros(sample_shape, device=dev, dtype=torch.float32) if needs_noise else None
        self.coef = torch.zeros(8, device=dev, dtype=torch.float32)
        self._coef_rows = {}
        self.op = scheduler.operator
        self._ir_host = self._ir_dev = None
        if hasattr(self.op, "generate_impulse_response"):  # dereverberation: IR redrawn on the host every step
            # a small ring of pinned staging rows: the host may run several steps ahead of the device, and a row must
            # not be rewritten before its asynchronous upload has executed
            self._ir_host = [torch.zeros(self.op.ir_length, dtype=torch.float32).pin_memory() for _ in range(4)]
            self._ir_event = [None] * 4
            self._ir_turn = 0
            self._ir_dev = torch.zeros(self.op.ir_length, device=dev, dtype=torch.float32)
            self.op.static_ir = self._ir_dev
        t0 = int(scheduler.timesteps[0]) if warmup_timestep is None else int(warmup_timestep)
        # the warm-up inputs must not shift anybody's random streams: the step noise of a generator-less warm-up comes
        # from the device's default generator and a dereverberation impulse response from the global CPU generator;
        # both states are restored, so a seeded run that builds a graph (including the driver's restart sub-batch
        # graphs) stays aligned with the eager and the reference streams
        with torch.random.fork_rng(devices=[dev]):
            self._prepare(t0, None, None)
        scheduler._coef_dev = self.coef
        # warm-up on a side stream (allocator pools, lazy module init, cached measurement transform), then capture
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(2):
                self._run(t0)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        from . import _lib
        n0 = _lib.launch_count()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = self._run(t0)
        #: number of this library's kernels inside one replay (kernel nodes captured from dm_* calls)
        self.kernels_per_replay = _lib.launch_count() - n0
        # the graph reads the operator's cached transform(measurement) at a fixed address: hold it, so an eager call with
        # another measurement (which replaces the operator's one-entry cache) cannot release it under the graph
        self._keepalive = dict(getattr(self.op, "_ref_cache", None) or {})
        scheduler._coef_dev = None
        if self._ir_dev is not None:
            self.op.static_ir = None  # eager calls on the same operator keep drawing their own impulse responses

    def _run(self, t):
        return self.sched.step(self.e, t, self.x, _noise=self.z if self.z is not None else _NO_NOISE, **self.kw)

    def _prepare(self, timestep, generator, variance_noise):
        """host-side per-step work: RNG draws (reference order), coefficients, impulse response."""
        if self.z is not None:
            z = self.sched.draw_step_noise(self.eta, generator, variance_noise, self.e, out=self.z)
            if z.data_ptr() != self.z.data_ptr():  # torch drew it (single generator / variance_noise given)
                self.z.copy_(z, non_blocking=True)
        elif self.eta > 0:
            self.sched.draw_step_noise(self.eta, generator, variance_noise, self.e)  # DDIM: discarded draw
        key = (timestep, self.sched.num_inference_steps)  # t_prev, hence sqrt_p / dir_coef / std, depends on the step count
        row = self._coef_rows.get(key)
        if row is None:  # one pinned 32-byte row per timestep, built once (host-side scalar math is slow)
            row = self.sched.coef_vector(timestep, self.eta, self.n_clip).clone().pin_memory()
            self._coef_rows[key] = row
        self.coef.copy_(row, non_blocking=True)
        if self._ir_host is not None:
            ir = self.op.generate_impulse_response(ir_length=self.op.ir_length, decay_factor=self.op.decay_factor)
            self.op.last_ir = ir
            k = self._ir_turn = (self._ir_turn + 1) % len(self._ir_host)
            if self._ir_event[k] is not None:
                self._ir_event[k].synchronize()  # its previous upload (4 steps ago) has executed: normally long done
            self._ir_host[k].copy_(ir.reshape(-1))
            # a kernel reads the pinned row over PCIe: a copy-engine transfer here would queue behind the bulk latent
            # uploads / downloads of HostPipelinedStep and stall the compute stream at the start of every step
            from . import _lib
            _lib.call("dm_copy_f32", self._ir_dev.data_ptr(), self._ir_host[k].data_ptr(), self._ir_dev.numel(),
                      _lib.stream(self._ir_dev.device))
            ev = self._ir_event[k] = self._ir_event[k] or torch.cuda.Event()
            ev.record()

    def replay_in_place(self, timestep, generator=None, variance_noise=None, prepared_event=None):
        """Replay on whatever the caller has already put into the static inputs `self.x` / `self.e` (e.g. uploaded
        straight from pinned host memory); returns the static output object, valid until the next replay.
        `prepared_event`: recorded after the step's small host-to-device transfers (coefficients, impulse response) and
        before the graph launch."""
        self._prepare(int(timestep), generator, variance_noise)
        if prepared_event is not None:
            prepared_event.record()
        self.graph.replay()
        return self.out

    def __call__(self, model_output, timestep, sample, generator=None, variance_noise=None, _borrow=False, **ignored):
        self.x.copy_(sample, non_blocking=True)
        self.e.copy_(model_output, non_blocking=True)
        self._prepare(int(timestep), generator, variance_noise)
        self.graph.replay()
        o = self.out
        if _borrow or not self.clone:
            return o
        return InverseProblemSchedulerOutput(
            prev_sample=o.prev_sample.clone(), pred_original_sample=o.pred_original_sample.clone(),
            loss=o.loss.clone() if o.loss.is_cuda else torch.tensor([int(timestep)]),
            loss_per_clip=None if o.loss_per_clip is None else o.loss_per_clip.clone())


class _NoNoise:
    """sentinel: 'noise handled by the caller, and there is none' (eta == 0)."""


_NO_NOISE = _NoNoise()


def _default_eta(scheduler):
    import inspect
    return inspect.signature(scheduler.step).parameters["eta"].default


class HostPipelinedStep:
    """Guided steps whose latents live in PINNED HOST memory, with the PCIe copies overlapped with the step itself.

    The reference moves one clip at a time between host and device around every call (run.py:286-312).  Here three
    streams work on consecutive steps at once: an upload stream brings step i+1's `(model_output, sample)` into one of
    two device staging slots while the captured graph of step i replays on the compute stream, and a download stream
    returns step i's `prev_sample` and per-clip loss while step i+1 computes.  All ordering is by CUDA events; nothing
    synchronises with the host until `drain()`.

        pipe = HostPipelinedStep(graphed)
        pipe.prefetch(e_host[0], x_host[0])
        for i, t in enumerate(timesteps):
            if i + 1 < n: pipe.prefetch(e_host[i + 1], x_host[i + 1])
            pipe.step(t, prev_host[i], loss_host[i])
        pipe.drain()
    """

    def __init__(self, graphed, second=None):
        """graphed: a GraphedGuidedStep.  second (optional): another GraphedGuidedStep captured with the same arguments;
        with two graphs the uploads land directly in the static inputs of the graph that will consume them and the
        downloads read its static outputs -- no device-to-device staging copies at all."""
        self.g = [graphed] if second is None else [graphed, second]
        if second is not None and (second.x.shape != graphed.x.shape or second.x.dtype != graphed.x.dtype):
            raise ValueError("the two captured steps must have the same latent shape and dtype")
        dev = graphed.x.device
        self.dev = dev
        self.direct = second is not None
        self.up, self.down = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        if self.direct:
            self.x_st, self.e_st = [graphed.x, second.x], [graphed.e, second.e]
        else:
            mk = lambda: torch.empty_like(graphed.x)  # noqa: E731
            self.x_st, self.e_st, self.prev_st = [mk(), mk()], [mk(), mk()], [mk(), mk()]
            self.loss_st = [torch.zeros(graphed.B, device=dev, dtype=torch.float32) for _ in range(2)]
        ev = lambda: [torch.cuda.Event(), torch.cuda.Event()]  # noqa: E731
        self.uploaded, self.slot_free, self.computed, self.downloaded = ev(), ev(), ev(), ev()
        self.prepared = torch.cuda.Event()  # the last step's own small transfers are through (see prefetch)
        self.n_up = self.n_run = 0
        self.bytes_up = 2 * graphed.x.numel() * graphed.x.element_size()
        self.bytes_down = graphed.x.numel() * graphed.x.element_size() + graphed.B * 4

    def prefetch(self, model_output_host, sample_host, after=None):
        """Queue the upload of the NEXT step's inputs (at most two steps may be in flight).  `after`: an event the
        upload must not start before (bench.py uses it to keep the copy inside the timed region)."""
        if self.n_up - self.n_run >= 2:
            raise RuntimeError("HostPipelinedStep: two uploads already in flight; call step() first")
        k = self.n_up % 2
        if after is not None:
            self.up.wait_event(after)
        if self.direct and self.n_run > 0:
            # bulk uploads start only once the step issued last has its own few-KB transfers (coefficients, impulse
            # response) behind it: on the PCIe link those would otherwise wait tens of microseconds behind 4 MB of
            # latents at the very start of the step
            self.up.wait_event(self.prepared)
        if self.n_up >= 2:
            self.up.wait_event(self.slot_free[k])  # the step that last used this slot has consumed it
        with torch.cuda.stream(self.up):
            self.x_st[k].copy_(sample_host, non_blocking=True)
            self.e_st[k].copy_(model_output_host, non_blocking=True)
            self.uploaded[k].record(self.up)
        self.n_up += 1

    def step(self, timestep, prev_sample_host, loss_host=None, generator=None, variance_noise=None):
        """Run the step whose inputs were prefetched first; its results land in the given pinned host tensors once
        `drain()` (or the second-next `step`) has returned."""
        if self.n_run >= self.n_up:
            raise RuntimeError("HostPipelinedStep: step() without a prefetched input")
        k = self.n_run % 2
        cur = torch.cuda.current_stream(self.dev)
        cur.wait_event(self.uploaded[k])
        if self.n_run >= 2:
            cur.wait_event(self.downloaded[k])  # this slot's previous results have reached the host
        if self.direct:
            out = self.g[k].replay_in_place(timestep, generator=generator, variance_noise=variance_noise,
                                            prepared_event=self.prepared)
            prev_src = out.prev_sample
            loss_src = None if out.loss_per_clip is None else out.loss_per_clip.reshape(-1)
        else:
            out = self.g[0](self.e_st[k], timestep, self.x_st[k], generator=generator, variance_noise=variance_noise,
                            _borrow=True)
            self.prev_st[k].copy_(out.prev_sample, non_blocking=True)
            prev_src, loss_src = self.prev_st[k], None
            if loss_host is not None and out.loss_per_clip is not None:
                self.loss_st[k].copy_(out.loss_per_clip.reshape(-1), non_blocking=True)
                loss_src = self.loss_st[k]
        self.slot_free[k].record(cur)  # the inputs of this slot have been consumed
        self.computed[k].record(cur)
        self.down.wait_event(self.computed[k])
        with torch.cuda.stream(self.down):
            prev_sample_host.copy_(prev_src, non_blocking=True)
            if loss_host is not None and loss_src is not None:
                loss_host.copy_(loss_src, non_blocking=True)
            self.downloaded[k].record(self.down)
        if self.n_run >= 1:
            cur.wait_event(self.downloaded[1 - k])  # step i returns only after step i-1's results are on the host
        self.n_run += 1

    def drain(self):
        """Make the current stream wait for every queued download (call before reading the host outputs)."""
        cur = torch.cuda.current_stream(self.dev)
        for k in range(2):
            if self.n_run > k:
                cur.wait_event(self.downloaded[k])
