"""world_size-2 gloo tests (CPU) of the N>1 host logic: clip sharding (no collective on the guided path) and the single
all-reduce of the packed FAD moments [n | sum x | upper triangle of sum x x^T] (fadtk/utils.py:19-46 equivalence)."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from diffmusic_b200 import parallel
from oracle import fad as ofad


def test_shard_indices_partition():
    for n, w in ((16, 2), (128, 8), (7, 4), (3, 8)):
        seen = sorted(i for r in range(w) for i in parallel.shard_indices(n, r, w))
        assert seen == list(range(n))
        assert all(i % w == r for r in range(w) for i in parallel.shard_indices(n, r, w))


def _worker(rank, world, port, d, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from diffmusic_b200.fad import EmbeddingMoments
    rng = np.random.default_rng(123)
    files = [(rng.standard_normal((10 + 3 * i, d)) * 0.5 + 0.2).astype(np.float16) for i in range(9)]
    mine = [files[i] for i in parallel.shard_indices(len(files), rank, world)]
    mom = EmbeddingMoments(d, device="cpu")
    # raw moments of this rank's shard (on the GPU box dm_fad_moments produces exactly this vector)
    for a in mine:
        a64 = a.astype(np.float64)
        mom.acc[0] += a64.shape[0]
        mom.acc[1:1 + d] += torch.from_numpy(a64.sum(0))
        mom.acc[1 + d:] += torch.from_numpy((a64.T @ a64).ravel())
    assert mom.packed().numel() == 1 + d + d * (d + 1) // 2   # only the upper triangle of sum x x^T travels
    mom.all_reduce()
    n = mom.count()
    acc = mom.moments()
    sx = acc[1:1 + d].numpy()
    up = np.triu(acc[1 + d:].numpy().reshape(d, d))
    sxx = up + np.triu(up, 1).T
    mu, cov = ofad.moments_to_stats(n, sx, sxx)
    want_mu, want_cov = ofad.embd_statistics_online([f.astype(np.float64) for f in files])
    ok = (n == sum(f.shape[0] for f in files) and np.allclose(mu, want_mu, rtol=1e-10, atol=1e-12)
          and np.allclose(cov, want_cov, rtol=1e-8, atol=1e-10))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_fad_allreduce_equals_chan_merge_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 48, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def _driver_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tests.test_driver import Predictor, gens, sampler
    meas = torch.arange(5, dtype=torch.float32).reshape(5, 1) * 0.1
    ids, out, table = parallel.run_sharded(sampler(Predictor(poison=[3]), eta=1.0), meas, gens(range(5)), gather=True)
    q.put((rank, ids, out.latents, table))
    dist.destroy_process_group()


def test_batched_driver_shards_clips_world2():
    """5 clips over 2 ranks (clip i -> rank i mod 2), no collective on the path: every clip equals the single-process batch,
    including the clip that restarts; the gathered loss table covers all clips on every rank."""
    from tests.test_driver import Predictor, gens, sampler
    meas = torch.arange(5, dtype=torch.float32).reshape(5, 1) * 0.1
    whole = sampler(Predictor(poison=[3]), eta=1.0)(meas, gens(range(5)))
    assert whole.restarts == [0, 0, 0, 1, 0]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_driver_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in procs), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
    assert [r[1] for r in res] == [[0, 2, 4], [1, 3]]
    for _, ids, lat, table in res:
        for k, i in enumerate(ids):
            assert torch.equal(lat[k], whole.latents[i])
        assert sorted(table) == [0, 1, 2, 3, 4]
        assert [table[i][1] for i in range(5)] == whole.restarts
        assert all(abs(table[i][0] - float(whole.loss[i])) < 1e-6 for i in range(5))
