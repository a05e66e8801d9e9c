"""2-GPU probe of the peer-memory FAD exchange (torchrun --nproc-per-node 2 tools/peer_probe.py): prints every stage so a
failure of the CUDA-IPC mapping or of the flag protocol is visible immediately."""
import faulthandler
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
faulthandler.dump_traceback_later(40, exit=True)


def log(*a):
    print(f"[rank {os.environ.get('RANK')}] {time.time() % 1000:8.3f}", *a, flush=True)


rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("gloo")
log("pg up; can_access_peer", [torch.cuda.can_device_access_peer(local, j) for j in range(world) if j != local])
from diffmusic_b200 import fad  # noqa: E402

d = 768
try:
    mom = fad.EmbeddingMoments(d, device=f"cuda:{local}", exchange="peer")
    torch.cuda.synchronize()
    log("peer group up", [hex(p) for p in mom.peers.ptrs])
    # stage 1: a kernel of this library on the local device reads the peer's buffer (dm_copy_f32) and writes its flag pad
    from diffmusic_b200 import _lib
    other = (rank + 1) % world
    mom.peers.buf[:8] = float(rank + 1)
    torch.cuda.synchronize()
    dist.barrier()
    got = torch.zeros(16, device=f"cuda:{local}", dtype=torch.float32)
    _lib.call("dm_copy_f32", got.data_ptr(), mom.peers.ptrs[other], 16, _lib.stream())
    torch.cuda.synchronize()
    log("kernel read of peer memory ok", got.view(torch.float64)[:2].tolist())
    mine = torch.full((2,), 7.0, device=f"cuda:{local}", dtype=torch.float32)
    _lib.call("dm_copy_f32", mom.peers.flag_ptrs[other] + 4 * 36, mine.data_ptr(), 2, _lib.stream())
    torch.cuda.synchronize()
    dist.barrier()
    log("kernel write to peer memory ok; my pad words 36..37 =", mom.peers.flags[36:38].view(torch.float32).tolist())
    mom.peers.buf[:8] = 0.0
    mom.peers.flags[36:38] = 0
    torch.cuda.synchronize()
    dist.barrier()
    for rnd in range(3):
        if rnd:
            mom.reset()
        g = torch.Generator().manual_seed(rnd * 10 + rank)
        x = (torch.randn(4990, d, generator=g) * 0.5 + 0.2).half()
        mom.update(x)
        torch.cuda.synchronize()
        log("moments done round", rnd)
        mom.all_reduce()
        torch.cuda.synchronize()
        log("all_reduce done; n =", mom.count())
        mu, cov = mom.finalize()
        torch.cuda.synchronize()
        log("finalize done", float(mu[0]), float(cov[0, 0]))
    s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    for it in range(3):
        mom.reset()
        mom.update(x)
        s.record()
        mom.all_reduce()
        t.record()
        torch.cuda.synchronize()
        log("exchange ms", s.elapsed_time(t))
    mom.close()
    log("closed")
except Exception:
    import traceback
    log("EXCEPTION\n" + traceback.format_exc())
dist.destroy_process_group()
