"""2-GPU tests of the one exchange on the path (run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`;
skipped on a single-GPU box): the FAD moment all-reduce over NVLink peer memory (csrc/fad_exchange.cu) and its NCCL
variant `dm_fad_allreduce` on a communicator created here, both against float64 NumPy on the full data
(fadtk/utils.py:19-46: the pairwise merge over files == the sum of raw moments over ranks)."""
import ctypes as C
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _data(rank_of, world, d, rounds):
    """per round: 6 seeded fp16 blocks; block i belongs to rank i mod world"""
    out = []
    for r in range(rounds):
        rng = np.random.default_rng(500 + r)
        out.append([(rng.standard_normal((300 + 40 * i, d)) * 0.5 + 0.2).astype(np.float16) for i in range(6)])
    return out


def _peer_worker(rank, world, port, d, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        torch.cuda.set_device(rank)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from diffmusic_b200 import _lib, fad
        rounds = _data(None, world, d, 3)
        ok = True
        one = fad.EmbeddingMoments(d, device=f"cuda:{rank}", exchange="peer_oneshot")  # A/B variant: same sums, bit for bit
        mom = fad.EmbeddingMoments(d, device=f"cuda:{rank}", exchange="peer")
        for i, b in enumerate(rounds[0]):
            if i % world == rank:
                one.update(torch.from_numpy(b))
        one.all_reduce()
        mu1, cov1 = one.finalize()
        for r, blocks in enumerate(rounds):
            if r > 0:
                mom.reset()
                mom.reset()  # any number of resets between two exchanges (round numbers count exchanges)
            for i, b in enumerate(blocks):
                if i % world == rank:
                    mom.update(torch.from_numpy(b))
            local = mom.acc.clone()
            mom.all_reduce()
            mu, cov = mom.finalize()       # its kernel waits until every rank's rows have landed
            got = mom.moments().clone()
            torch.cuda.synchronize()
            # (1) the sum over ranks, in rank order, of the ranks' own accumulators: bit-exact on the exchanged part
            parts = [torch.zeros_like(local).cpu() for _ in range(world)]
            dist.all_gather(parts, local.cpu())
            want = parts[0].clone()
            for p in parts[1:]:
                want += p
            up = torch.triu(torch.ones(d, d, dtype=torch.bool)).reshape(-1)
            ok &= bool(torch.equal(got.cpu()[:1 + d], want[:1 + d]))
            ok &= bool(torch.equal(got.cpu()[1 + d:][up], want[1 + d:][up]))
            # (2) the statistics of ALL blocks against float64 NumPy
            A = np.concatenate(blocks).astype(np.float64)
            emu = np.linalg.norm(mu.cpu().numpy() - A.mean(0)) / np.linalg.norm(A.mean(0))
            ecov = np.linalg.norm(cov.cpu().numpy() - np.cov(A, rowvar=False)) / np.linalg.norm(np.cov(A, rowvar=False))
            ok &= bool(mom.count() == A.shape[0] and emu < 1e-6 and ecov < 1e-5)
            if r == 0:
                ok &= bool(torch.equal(mu, mu1) and torch.equal(cov, cov1))
        # NCCL variant on a communicator of our own (ncclComm_t handed over as a raw pointer)
        import glob
        import torch as _t
        cands = glob.glob(os.path.join(os.path.dirname(_t.__file__), "..", "nvidia", "nccl", "lib", "libnccl.so*"))
        nccl = C.CDLL(cands[0] if cands else "libnccl.so.2", mode=C.RTLD_GLOBAL)
        uid = (C.c_char * 128)()
        if rank == 0:
            assert nccl.ncclGetUniqueId(C.byref(uid)) == 0
        box = [bytes(uid)]
        dist.broadcast_object_list(box, src=0)
        uid = (C.c_char * 128).from_buffer_copy(box[0])
        comm = C.c_void_p()

        class UID(C.Structure):
            _fields_ = [("internal", C.c_char * 128)]

        nccl.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, UID, C.c_int]
        u = UID()
        C.memmove(C.byref(u), uid, 128)
        assert nccl.ncclCommInitRank(C.byref(comm), world, u, rank) == 0
        blocks = rounds[0]
        m2 = fad.EmbeddingMoments(d, device=f"cuda:{rank}")
        for i, b in enumerate(blocks):
            if i % world == rank:
                m2.update(torch.from_numpy(b))
        work = torch.empty(int(_lib.load().dm_fad_packed_doubles(d)), device=f"cuda:{rank}", dtype=torch.float64)
        _lib.call("dm_fad_allreduce", comm, m2.acc.data_ptr(), d, work.data_ptr(), _lib.stream())
        mu2, cov2 = m2.finalize()
        torch.cuda.synchronize()
        A = np.concatenate(blocks).astype(np.float64)
        ok &= bool(np.linalg.norm(mu2.cpu().numpy() - A.mean(0)) / np.linalg.norm(A.mean(0)) < 1e-6)
        ok &= bool(np.linalg.norm(cov2.cpu().numpy() - np.cov(A, rowvar=False)) / np.linalg.norm(np.cov(A, rowvar=False))
                   < 1e-5)
        nccl.ncclCommDestroy.argtypes = [C.c_void_p]
        nccl.ncclCommDestroy(comm)
        mom.close()
        one.close()
        q.put((rank, bool(ok), ""))
        dist.destroy_process_group()
    except Exception as exc:  # surface the failure in the parent
        import traceback
        q.put((rank, False, traceback.format_exc()))


@pytest.mark.parametrize("d", [128, 768])
def test_fad_exchange_over_peer_memory_and_nccl(d):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 30500 + (os.getpid() % 2000) + d % 7
    procs = [ctx.Process(target=_peer_worker, args=(r, world, port, d, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = []
    try:
        for _ in procs:
            res.append(q.get(timeout=120))
            if not res[-1][1]:
                break  # a failed rank leaves its peer waiting in a collective: stop here
    finally:
        for p in procs:
            p.join(timeout=5 if res and not res[-1][1] else 60)
            if p.is_alive():
                p.terminate()
    res.sort()
    assert [(r, ok) for r, ok, _ in res] == [(0, True), (1, True)], [m for _, _, m in res]
