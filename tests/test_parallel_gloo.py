"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: particle sharding and the
single flat all-reduce of guide gradients (bayesair_b200/parallel.py)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bayesair_b200.parallel import allreduce_gradients, shard_particles


def test_shards_cover_every_particle_once():
    for P in (1, 7, 4096, 4097):
        for W in (1, 2, 3, 8):
            seen = []
            for r in range(W):
                s, c = shard_particles(P, r, W)
                seen += list(range(s, s + c))
            assert seen == list(range(P))
            sizes = [shard_particles(P, r, W)[1] for r in range(W)]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    from bayesair_b200.objective import MeanFieldGuide
    guide = MeanFieldGuide(12)
    # every rank evaluates its own shard of particles of the same objective
    start, count = shard_particles(10, rank, world)
    torch.manual_seed(100)
    eps = torch.randn(10, 12)[start:start + count]
    z = guide.loc + torch.exp(guide.log_scale) * eps
    loss = (z ** 2).sum() / 10.0
    loss.backward()
    extra = allreduce_gradients(guide.parameters(), extra=loss.detach().reshape(1))
    out.put((rank, guide.loc.grad.clone(), guide.log_scale.grad.clone(), extra.clone()))
    dist.barrier()
    dist.destroy_process_group()


def test_allreduce_matches_single_process():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process reference over all 10 particles
    from bayesair_b200.objective import MeanFieldGuide
    torch.manual_seed(0)
    guide = MeanFieldGuide(12)
    torch.manual_seed(100)
    eps = torch.randn(10, 12)
    z = guide.loc + torch.exp(guide.log_scale) * eps
    loss = (z ** 2).sum() / 10.0
    loss.backward()
    for _, gl, gs, extra in res:
        assert torch.allclose(gl, guide.loc.grad, atol=1e-6)
        assert torch.allclose(gs, guide.log_scale.grad, atol=1e-6)
        assert torch.allclose(extra, loss.detach().reshape(1), atol=1e-6)
