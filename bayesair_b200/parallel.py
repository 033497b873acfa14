"""Multi-GPU plumbing: shard particles across ranks, all-reduce guide gradients.

Particle-days are independent given a particle's latents (the reference re-initialises
every airport per day, ``bayes_air/model.py:126-155``), so the simulation itself needs no
collective: rank ``r`` runs a contiguous slice of the particles over all days.  The only
exchange per optimisation step is one flat ``all_reduce(SUM)`` of the guide-parameter
gradients (the reference has no distributed code at all; its only parallelism is one
process per seed, ``scripts/wn/experiments.sh:3-13``).  One process per GPU, NCCL over
NVLink/NVSwitch in production; the same code runs over gloo on CPU in the tests.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_particles(n_particles: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous, balanced slice ``(start, count)`` of ``n_particles`` for ``rank``."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, extra = divmod(n_particles, world_size)
    count = base + (1 if rank < extra else 0)
    start = rank * base + min(rank, extra)
    return start, count


def shard_particle_days(n_particles: int, n_days: int, rank: int, world_size: int
                        ) -> Tuple[int, int, int, int]:
    """2-D decomposition ``(particle_start, particle_count, day_start, day_count)`` of the
    (particle, day) grid for ``rank``.  Particles are split first; only when there are more
    ranks than particles can usefully fill (fewer than one particle per rank, or a particle
    count that does not divide over the ranks while the day count does) are the DAYS split as
    well: ``world_size = Gp x Gd`` with the largest ``Gp`` that divides ``world_size`` and is
    at most ``n_particles``.  Ranks of one particle shard share the guide draws (same seed,
    same ``particle_offset``) and each accounts for ``day_count / n_days`` of the per-particle
    terms (``prior_weight`` of ``ba_mf_backward``)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    gp = max(g for g in range(1, world_size + 1) if world_size % g == 0 and g <= max(1, n_particles))
    gd = world_size // gp
    if gd > n_days:
        raise ValueError(f"{world_size} ranks cannot be filled by {n_particles} particles x {n_days} days")
    rp, rd = divmod(rank, gd)
    p0, pc = shard_particles(n_particles, rp, gp)
    d0, dc = shard_particles(n_days, rd, gd)
    return p0, pc, d0, dc


def allreduce_gradients(params: Iterable[torch.nn.Parameter], extra: Optional[torch.Tensor] = None,
                        group=None, average: bool = False) -> Optional[torch.Tensor]:
    """Sums ``p.grad`` of every parameter (and ``extra``, e.g. the scalar loss or the three
    noise-scale gradients) over all ranks with ONE collective on one flat fp32 buffer."""
    # EVERY parameter takes part, in the caller's order (a missing gradient counts as zero):
    # the flat layout is then identical on all ranks whatever each rank's graph touched
    plist: List[torch.nn.Parameter] = list(params)
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return extra
    if not plist and extra is None:
        return None
    ref = plist[0] if plist else extra
    chunks = [(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).to(torch.float32)
              for p in plist]
    if extra is not None:
        chunks.append(extra.reshape(-1).to(torch.float32))
    flat = torch.cat(chunks)
    with torch.cuda.nvtx.range("ba.allreduce") if ref.is_cuda else _no_range():
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= dist.get_world_size(group)
    off = 0
    for p in plist:
        n = p.numel()
        g = flat[off:off + n].view_as(p).to(p.dtype)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        off += n
    if extra is not None:
        return flat[off:].view_as(extra).to(extra.dtype).clone()
    return None


class _no_range:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
