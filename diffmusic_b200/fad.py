"""FAD embedding statistics on the GPU: mean / covariance of (N, d) fp16 embeddings (fadtk/fad.py:41-47) and the
multi-file / multi-rank merge (fadtk/utils.py:13-46).

Every shard (file, clip batch or rank) contributes raw float64 moments  acc = [n | sum x | sum x x^T]  through
dm_fad_moments; shards add, ranks add with ONE all-reduce (NCCL over NVLink when torch.distributed is initialised with
the nccl backend), and  mu = sx / n,  cov = (sxx - n mu mu^T) / (n - 1)  is the Chan merge the reference computes
pairwise on the host.
"""
from __future__ import annotations

import torch

from . import _lib


class EmbeddingMoments:
    """Accumulator of raw moments for one embedding model of width d.

    exchange: how `all_reduce` sums the moments over ranks (only the upper triangle of sum x x^T travels either way):
      "peer"  -- one kernel per rank over NVLink (CUDA-IPC mapped memory, no collective library, no host
                 synchronisation; csrc/fad_exchange.cu): rank r reduces the triangle rows r, r + W, ... from every peer
                 and pushes them into every rank's sum buffer.  The accumulator lives in peer-visible memory and is
                 reused round after round: call `reset()` to start the next one.  "peer_oneshot": every rank reads the
                 whole triangle of every peer instead (A/B).
      "nccl"  -- torch.distributed all_reduce of the packed triangle (any backend; what the gloo tests exercise).
      None    -- "nccl" when torch.distributed is initialised with more than one rank, else nothing to do."""

    ENGINES = {"auto": 0, "simt": 1, "tcgen05": 2, "tcgen05_pair": 3}

    def __init__(self, d, device=None, engine="auto", exchange=None, group=None):
        self.d = int(d)
        self.engine = self.ENGINES[engine]
        self.exchange = exchange
        self.group = group
        # the accumulator may live on the CPU (gloo tests of the exchange step); update()/finalize() need CUDA
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        n = 1 + self.d + self.d * self.d
        self.peers = None
        self._reduced = None
        self._wait_done = False
        self._round = 0
        if exchange in ("peer", "peer_oneshot", "peer_push"):
            from . import parallel
            # one peer-visible allocation per rank: [accumulator | sum buffer | flag pad]
            self.peers = parallel.PeerGroup(2 * n, int(_lib.load().dm_fad_flag_words()), device=self.device,
                                            group=group)
            self.acc = self.peers.buf[:n]
            self._sum = self.peers.buf[n:]
            import ctypes as C
            W = self.peers.world
            self._acc_ptrs = (C.c_void_p * W)(*self.peers.ptrs)
            self._sum_ptrs = (C.c_void_p * W)(*[q + 8 * n for q in self.peers.ptrs])
            self._flag_ptrs = (C.c_void_p * W)(*self.peers.flag_ptrs)
            self._wait_done = False
            self.reset()
        else:
            self.acc = torch.zeros(n, device=self.device, dtype=torch.float64)

    def reset(self):
        """start a new round: clear the accumulator (peer mode: once every peer has finished reading the last round)"""
        self._reduced = None
        self._wait_done = False
        if self.peers is not None:
            # wait for the peers' DONE of the last exchange this accumulator took part in (round numbers count
            # exchanges, so any number of resets between two exchanges is fine)
            _lib.call("dm_fad_reset_shared", self.acc.data_ptr(), self.d, self.peers.flags.data_ptr(), self.peers.world,
                      self._round, _lib.stream())
        else:
            self.acc.zero_()
        return self

    def update(self, embd):
        """embd: (n_frames, d) fp16 (what fadtk caches, model_loader.py:46-48) or any float dtype (cast to fp16 only
        if it already is fp16-representable is the caller's business: fp32 input is rounded to fp16 like the cache)."""
        if embd.dim() != 2 or embd.shape[1] != self.d:
            raise ValueError(f"expected (n, {self.d}) embeddings, got {tuple(embd.shape)}")
        if self.device.type != "cuda":
            raise _lib.DiffMusicB200Error("dm_fad_moments needs a CUDA accumulator (no CPU fallback)")
        x = embd.to(device=self.device, dtype=torch.float16).contiguous()
        if x.shape[0] == 0:
            return self
        _lib.call("dm_fad_moments_ex", x.data_ptr(), x.shape[0], self.d, self.acc.data_ptr(), self.engine,
                  _lib.stream())
        return self

    def all_reduce(self, group=None):
        """Sum the moments over ranks: the only exchange on this path, 1 + d + d (d + 1) / 2 float64 per rank."""
        import torch.distributed as dist
        group = group if group is not None else self.group
        if self.peers is not None:
            self._round += 1
            if self.exchange == "peer_oneshot" or (self.exchange == "peer" and self.peers.world <= 2):
                # two ranks: reading the peer's triangle once moves the same bytes as reduce-and-push, in one phase
                _lib.call("dm_fad_allreduce_peers", self._acc_ptrs, self._flag_ptrs, self.peers.world, self.peers.rank,
                          self.d, self._round, self._sum.data_ptr(), _lib.stream())
            else:  # reduce my rows, push them to everybody; complete once every rank has raised DONE (see finalize)
                _lib.call("dm_fad_allreduce_push", self._acc_ptrs, self._sum_ptrs, self._flag_ptrs, self.peers.world,
                          self.peers.rank, self.d, self._round, _lib.stream())
                self._wait_done = True
            self._reduced = self._sum
            return self
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            packed = self.packed()
            dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
            self._reduced = self.unpacked(packed)
        return self

    # packed = [n | sum x | rows i of sum x x^T from column i on]: what travels between ranks
    def packed(self):
        d = self.d
        if self.acc.is_cuda:
            out = torch.empty(int(_lib.load().dm_fad_packed_doubles(d)), device=self.device, dtype=torch.float64)
            _lib.call("dm_fad_pack_tri", self.acc.data_ptr(), d, out.data_ptr(), _lib.stream())
            return out
        iu = torch.triu_indices(d, d)
        return torch.cat([self.acc[:1 + d], self.acc[1 + d:].reshape(d, d)[iu[0], iu[1]]])

    def unpacked(self, packed):
        """accumulator with (at least) the upper triangle of sum x x^T filled from `packed`"""
        d = self.d
        out = torch.zeros_like(self.acc)
        if self.acc.is_cuda:
            _lib.call("dm_fad_unpack_tri", packed.data_ptr(), d, out.data_ptr(), _lib.stream())
            return out
        iu = torch.triu_indices(d, d)
        out[:1 + d] = packed[:1 + d]
        out[1 + d:].reshape(d, d)[iu[0], iu[1]] = packed[1 + d:]
        return out

    def moments(self):
        """the (all-reduced, if all_reduce ran) accumulator; sum x x^T is guaranteed on the upper triangle only.  In
        "peer" mode the sum is complete on the stream only after `finalize()` (whose kernel waits for every rank's
        pushed rows); n and sum x are this rank's own work and valid right after `all_reduce()`."""
        return self._reduced if self._reduced is not None else self.acc

    def count(self):
        return int(round(float(self.moments()[0].item())))

    def finalize(self):
        """(mu (d,), cov (d, d)) float64 on the device; cov is zeros when fewer than 2 frames (fadtk/utils.py:42-46)."""
        if self.device.type != "cuda":
            raise _lib.DiffMusicB200Error("dm_fad_finalize needs a CUDA accumulator (no CPU fallback)")
        mu = torch.empty(self.d, device=self.device, dtype=torch.float64)
        cov = torch.empty((self.d, self.d), device=self.device, dtype=torch.float64)
        if self.peers is not None and self._reduced is not None and self._wait_done:
            _lib.call("dm_fad_finalize_shared", self._sum.data_ptr(), self.d, self.peers.flags.data_ptr(),
                      self.peers.world, self._round, mu.data_ptr(), cov.data_ptr(), _lib.stream())
        else:
            _lib.call("dm_fad_finalize_sym", self.moments().data_ptr(), self.d, mu.data_ptr(), cov.data_ptr(),
                      _lib.stream())
        return mu, cov

    def close(self):
        if self.peers is not None:
            self.peers.close()


def calc_embd_statistics(embd_lst):
    """fadtk/fad.py:41-47 on the GPU: (mean, cov) of one (n, d) block, returned as float64 NumPy arrays."""
    if embd_lst.shape[0] < 2:
        raise AssertionError(f"FAD requires at least two embedding window frames, you have {tuple(embd_lst.shape)}.")
    t = torch.as_tensor(embd_lst)
    mu, cov = EmbeddingMoments(t.shape[1]).update(t).finalize()
    return _mean_like_numpy(mu, t.dtype).cpu().numpy(), cov.cpu().numpy()


def _mean_like_numpy(mu, in_dtype):
    """np.mean keeps a floating input's dtype: for the fp16 arrays fadtk caches (model_loader.py:46-48) the reference's
    mean is an fp16 array while np.cov is float64.  The exact float64 mean is rounded to that dtype
    once (numpy rounds a float32-accumulated mean: the two can differ by one fp16 ulp in rare elements)."""
    if in_dtype in (torch.float16, torch.float32):
        return mu.to(in_dtype)
    return mu


def calculate_embd_statistics_online(arrays, group=None):
    """fadtk/utils.py:19-46 with in-memory blocks (or .npy paths) instead of a file list; all-reduced over `group`
    when torch.distributed is initialised, so each rank may pass only its own shard."""
    import numpy as np
    if len(arrays) == 0:
        raise AssertionError("No files provided")
    first = np.load(arrays[0]) if isinstance(arrays[0], (str, bytes)) or hasattr(arrays[0], "__fspath__") else arrays[0]
    mom = EmbeddingMoments(first.shape[-1])
    for a in arrays:
        if isinstance(a, (str, bytes)) or hasattr(a, "__fspath__"):
            a = np.load(a)
        mom.update(torch.as_tensor(a))
    mom.all_reduce(group)
    mu, cov = mom.finalize()
    return mu.cpu().numpy(), cov.cpu().numpy()


# ---------------------------------------------------------------------------------------------------------------------
# Frechet distance and FAD-inf (fadtk/fad.py:50-119, 303-350)
JACOBI_MAX_SWEEPS = 40        # cap of the retry
JACOBI_FIRST_SWEEPS = 16      # sweeps enqueued up front (one graph replay of d - 1 rounds each), no-ops once converged
JACOBI_TOL = 1e-14


def _as_dev_f64(a, device):
    return torch.as_tensor(a).to(device=device, dtype=torch.float64).contiguous()


def frechet_distance_device(mu1, cov1, mu2, cov2, max_sweeps=JACOBI_FIRST_SWEEPS):
    """Device tensors in, (4,) float64 device tensor out: [d^2, tr sqrt(C1 C2), sweeps of the two Jacobi solves].
    No host synchronisation (dm_frechet_distance).  A solve that used all `max_sweeps` sweeps may not have converged:
    callers that can read the result check out[2:4] and retry with JACOBI_MAX_SWEEPS (`_frechet_checked`)."""
    _lib.require_cuda(mu1, cov1, mu2, cov2)
    d = mu1.numel()
    if mu2.numel() != d:
        raise AssertionError(f"Training and test mean vectors have different lengths ({tuple(mu1.shape)} vs "
                             f"{tuple(mu2.shape)})")
    if tuple(cov1.shape) != (d, d) or tuple(cov2.shape) != (d, d):
        raise AssertionError(f"Training and test covariances have different dimensions ({tuple(cov1.shape)} vs "
                             f"{tuple(cov2.shape)})")
    work = torch.empty(int(_lib.load().dm_frechet_workspace_doubles(d)), device=mu1.device, dtype=torch.float64)
    out = torch.empty(4, device=mu1.device, dtype=torch.float64)
    _lib.call("dm_frechet_distance", mu1.data_ptr(), cov1.data_ptr(), mu2.data_ptr(), cov2.data_ptr(), d,
              int(max_sweeps), JACOBI_TOL, work.data_ptr(), out.data_ptr(), _lib.stream())
    return out


def _frechet_checked(outs, args):
    """outs: list of (4,) device results of frechet_distance_device(*args[i]) with JACOBI_FIRST_SWEEPS; one host read
    for all of them, and a re-run with the full sweep budget for any solve that hit the cap."""
    host = torch.stack(outs).cpu()
    for i in range(len(outs)):
        if max(float(host[i, 2]), float(host[i, 3])) >= JACOBI_FIRST_SWEEPS:
            host[i] = frechet_distance_device(*args[i], max_sweeps=JACOBI_MAX_SWEEPS).cpu()
    return host


def calc_frechet_distance(mu1, cov1, mu2, cov2, eps=1e-6):
    """fadtk/fad.py:50-119 on the GPU.  Same value as the reference's eigenvalue method (sum of sqrt of eig(C1 C2)),
    obtained from two symmetric Jacobi eigen-solves (csrc/fad_frechet.cu); `eps` is accepted for signature parity -- the
    symmetric formulation has no singular-product failure mode to patch."""
    import numpy as np
    mu1, mu2 = np.atleast_1d(mu1), np.atleast_1d(mu2)
    cov1, cov2 = np.atleast_2d(cov1), np.atleast_2d(cov2)
    assert mu1.shape == mu2.shape, \
        f'Training and test mean vectors have different lengths ({mu1.shape} vs {mu2.shape})'
    assert cov1.shape == cov2.shape, \
        f'Training and test covariances have different dimensions ({cov1.shape} vs {cov2.shape})'
    if not torch.cuda.is_available():
        raise _lib.DiffMusicB200Error("dm_frechet_distance needs CUDA (no CPU fallback)")
    dev = torch.device("cuda", torch.cuda.current_device())
    args = (_as_dev_f64(mu1, dev), _as_dev_f64(cov1, dev), _as_dev_f64(mu2, dev), _as_dev_f64(cov2, dev))
    d2 = float(_frechet_checked([frechet_distance_device(*args)], [args])[0, 0])
    # The mean term in the reference's own arithmetic: `diff = mu1 - mu2; diff.dot(diff)` runs in the dtype NumPy
    # promotes the two means to (fadtk/fad.py:83, 112) -- fp16 when both come from fp16 embedding caches,
    # which rounds |mu1 - mu2|^2 to 11 bits.  d of these on the host; the trace terms stay on the device.
    diff = mu1 - mu2
    exact = (mu1.astype(np.float64) - mu2.astype(np.float64))
    return d2 - float(exact.dot(exact)) + float(diff.dot(diff))


class FADInfResults(tuple):
    """(score, slope, r2, points) -- fadtk/fad.py:34-38."""
    __slots__ = ()
    _fields = ("score", "slope", "r2", "points")

    def __new__(cls, score, slope, r2, points):
        return tuple.__new__(cls, (score, slope, r2, points))

    score = property(lambda s: s[0])
    slope = property(lambda s: s[1])
    r2 = property(lambda s: s[2])
    points = property(lambda s: s[3])


SCORE_INF_STREAMS = 4
_STREAMS = {}


def _side_streams(dev, k):
    pool = _STREAMS.setdefault(str(dev), [])
    while len(pool) < k:
        pool.append(torch.cuda.Stream(device=dev))
    return pool[:k]


def score_inf(mu_base, cov_base, embeds, steps=25, min_n=500):
    """fadtk/fad.py:303-350 with the statistics on the GPU: for `steps` sample sizes n between min_n and len(embeds),
    draw n rows with replacement (np.random.choice -- the reference's generator and draw order, so a seeded run picks
    the same rows), gather them on the device, take mean / covariance with the tcgen05 moment kernel and the Frechet
    distance against the baseline; FAD-inf is the intercept of the linear fit over 1/n."""
    import numpy as np
    if not torch.cuda.is_available():
        raise _lib.DiffMusicB200Error("score_inf needs CUDA (no CPU fallback)")
    dev = torch.device("cuda", torch.cuda.current_device())
    emb_dtype = torch.as_tensor(embeds).dtype
    x = torch.as_tensor(embeds).to(device=dev, dtype=torch.float16).contiguous()
    N, d = x.shape
    mu_b, cov_b = _as_dev_f64(np.atleast_1d(mu_base), dev), _as_dev_f64(np.atleast_2d(cov_base), dev)
    ns = [int(n) for n in np.linspace(min_n, N, steps)]
    base_dtype = np.atleast_1d(mu_base).dtype
    outs, args, fixes = [], [], []
    # The points are independent and a Jacobi solve is a chain of ~2-4 us rounds that leaves most of the GPU idle: they
    # are spread round-robin over a few side streams (rows are still drawn in the reference's order on the host).
    cur = torch.cuda.current_stream(dev)
    pool = _side_streams(dev, min(SCORE_INF_STREAMS, len(ns))) if SCORE_INF_STREAMS > 1 else [cur]
    for s in pool:
        if s is not cur:
            s.wait_stream(cur)  # x and the baseline statistics are ready
    for i, n in enumerate(ns):
        indices = np.random.choice(N, size=n, replace=True)
        with torch.cuda.stream(pool[i % len(pool)]):
            idx = torch.from_numpy(np.ascontiguousarray(indices, dtype=np.int64)).to(dev)
            sub = torch.empty((n, d), device=dev, dtype=torch.float16)
            _lib.call("dm_fad_gather_rows", x.data_ptr(), N, d, idx.data_ptr(), n, sub.data_ptr(), _lib.stream())
            mu, cov = EmbeddingMoments(d, device=dev).update(sub).finalize()
            # the reference takes np.mean of the fp16 rows (fad.py:334): an fp16-rounded mean enters the distance
            mu = _mean_like_numpy(mu, emb_dtype).to(torch.float64)
            args.append((mu_b, cov_b, mu, cov))
            outs.append(frechet_distance_device(*args[-1]))
            if base_dtype == np.float16 and emb_dtype == torch.float16:
                # both means are fp16 arrays in the reference: `diff.dot(diff)` is fp16 arithmetic (see
                # calc_frechet_distance); the same rounding here, on the device, without a host read
                diff16 = mu_b.to(torch.float16) - mu.to(torch.float16)
                quirk = (diff16.double() ** 2).sum().to(torch.float16).double()
                fixes.append(quirk - ((mu_b - mu) ** 2).sum())
            else:
                fixes.append(torch.zeros((), device=dev, dtype=torch.float64))
    for s in pool:
        if s is not cur:
            cur.wait_stream(s)
    fix = torch.stack(fixes)
    fad = _frechet_checked(outs, args)[:, 0].numpy() + fix.cpu().numpy()  # one synchronisation for the whole sweep
    results = [[n, float(s)] for n, s in zip(ns, fad)]
    ys = np.array(results)
    xs = 1 / np.array(ns)
    slope, intercept = np.polyfit(xs, ys[:, 1], 1)
    r2 = 1 - np.sum((ys[:, 1] - (slope * xs + intercept)) ** 2) / np.sum((ys[:, 1] - np.mean(ys[:, 1])) ** 2)
    return FADInfResults(score=intercept, slope=slope, r2=r2, points=results)


def patch_fadtk(fad_module=None, utils_module=None):
    """Point an imported `fadtk` at the GPU implementations: `fadtk.fad.calc_embd_statistics`,
    `fadtk.fad.calc_frechet_distance` (fad.py:41-47, 50-119) and `fadtk.utils.calculate_embd_statistics_online`
    (utils.py:19-46, which takes a list of .npy paths -- accepted here as well).  `fadtk` is a regular package, so unlike
    the `diffmusic` namespace it cannot be shadowed on sys.path; call this once after importing it:

        import fadtk.fad, fadtk.utils
        from diffmusic_b200.fad import patch_fadtk
        patch_fadtk(fadtk.fad, fadtk.utils)

    `FrechetAudioDistanceTK.score` / `score_inf` / `load_stats` then run their statistics and distances on the GPU (they
    look the functions up in the module namespace at call time).  Returns the names that were replaced."""
    done = []
    if fad_module is None:
        import fadtk.fad as fad_module  # noqa: PLC0415
    for name, fn in (("calc_embd_statistics", calc_embd_statistics), ("calc_frechet_distance", calc_frechet_distance)):
        if hasattr(fad_module, name):
            setattr(fad_module, name, fn)
            done.append(f"{fad_module.__name__}.{name}")
    if utils_module is None:
        try:
            import fadtk.utils as utils_module  # noqa: PLC0415
        except Exception:  # hypy_utils & co. may be missing; the fad module is what the scores go through
            utils_module = None
    if utils_module is not None and hasattr(utils_module, "calculate_embd_statistics_online"):
        utils_module.calculate_embd_statistics_online = calculate_embd_statistics_online
        done.append(f"{utils_module.__name__}.calculate_embd_statistics_online")
    if hasattr(fad_module, "calculate_embd_statistics_online"):  # `from .utils import *` copies it into fadtk.fad
        fad_module.calculate_embd_statistics_online = calculate_embd_statistics_online
        done.append(f"{fad_module.__name__}.calculate_embd_statistics_online")
    return done
