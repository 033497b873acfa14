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


def test_particle_day_grid_is_covered_once():
    from bayesair_b200.parallel import shard_particle_days
    for P, D, W in ((4096, 10, 8), (4, 30, 8), (1, 8, 8), (3, 10, 6), (7, 5, 1), (2, 2, 4)):
        seen = {}
        weight = {}
        for r in range(W):
            p0, pc, d0, dc = shard_particle_days(P, D, r, W)
            for p in range(p0, p0 + pc):
                weight[p] = weight.get(p, 0.0) + dc / D
                for d in range(d0, d0 + dc):
                    seen[(p, d)] = seen.get((p, d), 0) + 1
        assert len(seen) == P * D and set(seen.values()) == {1}, (P, D, W)
        # the ranks of one particle shard account for the whole prior between them
        assert all(abs(w - 1.0) < 1e-12 for w in weight.values())
    # plenty of particles: the days stay whole
    assert shard_particle_days(4096, 10, 3, 8) == (1536, 512, 0, 10)


def _flat_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bayesair_b200.calnf import FlatParameters
    from bayesair_b200.flows import ConditionalSplineFlow
    torch.manual_seed(0)                                  # identical weights on every rank
    flow = ConditionalSplineFlow(6, 2, hidden_features=(8, 8), transforms=2)
    for layer in flow.layers:
        torch.nn.init.normal_(layer.out.weight, std=0.2)
    fp = FlatParameters(flow)
    # every rank differentiates its own shard of the same sample set
    torch.manual_seed(100)
    z_all = torch.randn(8, 6)
    start, count = shard_particles(8, rank, world)
    loss = -flow(torch.tensor([1.0, 0.0])).log_prob(z_all[start:start + count]).sum() / 8.0
    loss.backward()
    fp.all_reduce()
    out.put((rank, fp.grad.clone(), fp.nbytes))
    dist.barrier()
    dist.destroy_process_group()


def test_flat_gradient_buffer_allreduce_matches_single_process():
    """The CalNF exchange (calnf.FlatParameters.all_reduce): one collective on the buffer
    autograd wrote; two ranks over half the samples each equal one process over all."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_flat_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from bayesair_b200.calnf import FlatParameters
    from bayesair_b200.flows import ConditionalSplineFlow
    torch.manual_seed(0)
    flow = ConditionalSplineFlow(6, 2, hidden_features=(8, 8), transforms=2)
    for layer in flow.layers:
        torch.nn.init.normal_(layer.out.weight, std=0.2)
    fp = FlatParameters(flow)
    torch.manual_seed(100)
    z_all = torch.randn(8, 6)
    (-flow(torch.tensor([1.0, 0.0])).log_prob(z_all).sum() / 8.0).backward()
    for _, g, nbytes in res:
        assert nbytes == fp.nbytes
        assert torch.allclose(g, fp.grad, atol=1e-5, rtol=1e-4)
