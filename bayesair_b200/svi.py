"""Fused SVI step for a mean-field guide (SURVEY.md §8e "fusion point").

One optimisation step of the reference's WN objective (``scripts/wn/train.py:301-326``:
``loss = -mean_p ELBO_p / flights``, ``ELBO = model log-prob - guide log-prob``) with a
diagonal-Normal guide over the ``2N + N^2`` log-latents, run entirely as library kernels:

    ba_mf_sample  ->  ba_forward  ->  ba_mf_backward  ->  all-reduce(flat)  ->  fused Adam

``ba_mf_backward`` contracts the particle-day gradient rows of the simulation with the
guide's (elementwise) Jacobian and writes ONE flat buffer ``[d loc | d log_scale | d scales |
loss]`` -- the NCCL send buffer -- so that no framework autograd, no ``torch.cat`` and no
per-parameter copies remain in the step.  Parameters live in one flat tensor whose ``.grad``
aliases that buffer; Adam runs as a single fused, capturable kernel; the Philox seed is a
device word, so the whole step replays as a CUDA graph; the per-particle-day status words
are OR-ed into a device flag that the host reads only when asked (``check()``).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import torch
import torch.distributed as dist

from .objective import n_latent
from .plan import SimulationPlan


def _nvtx(name: str):
    return torch.cuda.nvtx.range(name)


def _p(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


class FusedMeanFieldSVI:
    """``step()`` = one SVI step on this rank's ``n_particles`` (all days of the plan);
    gradients are summed over the process group, the loss is the global mean."""

    def __init__(self, plan: SimulationPlan, scales: torch.Tensor, n_particles: int, *,
                 init_loc: Optional[torch.Tensor] = None, init_scale: float = 0.05,
                 lr: float = 1e-3, seed: int = 0, group=None, use_graph: bool = True,
                 particle_offset: Optional[int] = None, n_particles_total: Optional[int] = None,
                 n_flights_total: Optional[int] = None, prior_weight: float = 1.0,
                 day_offset: int = 0):
        """``plan``: this rank's days.  By default the ranks shard the PARTICLES (rank r owns
        particles ``[r P, (r + 1) P)`` of ``world x P``, the plan holds every day).  With the
        days sharded as well (``parallel.shard_particle_days``), pass the shard explicitly:
        ``particle_offset`` (first particle of this rank's shard), ``n_particles_total``,
        ``n_flights_total`` (flights of ALL days), ``day_offset`` (first day of this rank's
        plan) and ``prior_weight`` = this rank's share of the days (the ranks of one particle
        shard draw identical latents)."""
        if plan.include_cancellations:
            raise ValueError("the fused mean-field step supports plans without cancellations")
        self.plan, self.P, self.group = plan, int(n_particles), group
        dev = plan.device
        N = plan.n_airports
        self.n = n_latent(N)
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        self.offset = self.rank * self.P if particle_offset is None else int(particle_offset)
        total_p = self.world * self.P if n_particles_total is None else int(n_particles_total)
        total_f = plan.n_flights_total if n_flights_total is None else int(n_flights_total)
        self.inv = 1.0 / (total_p * total_f)
        self.prior_weight = float(prior_weight)
        self.day_offset = int(day_offset)
        theta = torch.zeros(2 * self.n, device=dev)
        if init_loc is not None:
            theta[:self.n] = init_loc.to(dev)
        theta[self.n:] = math.log(init_scale)
        self.theta = theta.requires_grad_(True)          # loc | log_scale
        self.flat = torch.zeros(2 * self.n + 4, device=dev)
        self.theta.grad = self.flat[:2 * self.n]
        self.scales = scales.to(dev).float().contiguous()
        self.seed = torch.tensor([int(seed)], dtype=torch.int64, device=dev)
        self.status = torch.zeros((), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            self.la = torch.empty((self.P, 2, N), device=dev)
            self.lr_ = torch.zeros((self.P, plan.n_routes), device=dev)
            self.lp0 = torch.empty((self.P,), device=dev)
            self.work = torch.empty((plan._h.mf_work_bytes(self.P) // 4,), device=dev)
        self.opt = torch.optim.Adam([self.theta], lr=lr, fused=True, capturable=True)
        self._graph = None
        self._use_graph = bool(use_graph)
        self._steps = 0

    # ---- views ---------------------------------------------------------------------
    @property
    def loc(self):
        return self.theta.detach()[:self.n]

    @property
    def log_scale(self):
        return self.theta.detach()[self.n:]

    @property
    def loss(self):
        """Loss of the last step (device scalar; reading it synchronises)."""
        return self.flat[2 * self.n + 3]

    # ---- one step ----------------------------------------------------------------------
    def local_gradient(self, z_out: Optional[torch.Tensor] = None):
        """Sample -> simulate -> contract on this rank's particles: fills ``flat`` with this
        rank's share of ``[d loc | d log_scale | d scales | loss]`` (no collective, no
        optimiser).  ``z_out [P, n]`` receives the sampled log-latents (diagnostics)."""
        plan, h = self.plan, self.plan._h
        stream = C.c_void_p(torch.cuda.current_stream(plan.device).cuda_stream)
        h.mf_sample(self.P, _p(self.theta), 0, _p(self.seed), self.offset, _p(self.la),
                    _p(self.lr_), _p(self.lp0), _p(z_out), stream)
        out = plan.forward(self.la, self.lr_, self.scales, None, want_grad=True,
                           seed_tensor=self.seed, particle_offset=self.offset,
                           day_offset=self.day_offset)
        h.mf_backward(self.P, _p(self.theta), 0, _p(self.seed), self.offset, _p(out["tape"]),
                      _p(out["logp"]), _p(self.lp0), self.inv, self.prior_weight, _p(self.work),
                      _p(self.flat), stream)
        return out

    def _enqueue(self):
        with _nvtx("ba.svi.local_gradient"):
            out = self.local_gradient()
        self.status.bitwise_or_(out["status"].amax().to(torch.int32))
        if self.world > 1:
            with _nvtx("ba.svi.allreduce"):
                dist.all_reduce(self.flat, group=self.group)
        with _nvtx("ba.svi.adam"):
            self.opt.step()
        self.seed.add_(1)
        return out

    def step(self) -> torch.Tensor:
        """Enqueues one step; returns the loss as a device scalar (no host sync)."""
        dev = self.plan.device
        if not self._use_graph:
            with torch.cuda.device(dev):
                self._keep = self._enqueue()
        elif self._graph is None and self._steps < 3:
            with torch.cuda.device(dev):               # warm-up: allocations, NCCL setup
                self._keep = self._enqueue()
        elif self._graph is None:
            with torch.cuda.device(dev):
                g = torch.cuda.CUDAGraph()
                torch.cuda.synchronize(dev)
                try:
                    with torch.cuda.graph(g):
                        self._keep = self._enqueue()
                    self._graph = g
                    g.replay()
                except Exception:                       # capture not possible here: stay eager
                    self._use_graph = False
                    torch.cuda.synchronize(dev)
                    self._keep = self._enqueue()
        else:
            self._graph.replay()
        self._steps += 1
        return self.loss

    def check(self):
        """Raises if any particle-day of any step so far failed (synchronises)."""
        bits = int(self.status.item())
        if bits:
            SimulationPlan.check_status(torch.tensor([bits], dtype=torch.int32))
