"""Batched ELBO objective over the fused kernels.

Vectorised counterpart of the reference's WN objective closures
(``scripts/wn/train.py:228-326``): the guide proposes a ``[P, n_latent]`` sample of
log-latents laid out exactly like the reference's vector
(``[turnaround N | service N | (log n0 N | base-cancel N) | travel N x N]``,
``train.py:228-235,254-267``), ``map_to_latents`` is the tensor form of
``map_to_sample_sites`` (``train.py:244-297``), the model log-joint is
``prior(latents) + kernel log-prob`` and ``elbo = model_logprob - guide_logprob``
(``single_particle_elbo``, ``train.py:301-316``) for all P particles at once instead of a
Python loop over particles (``train.py:318-326``).  The priors are the ones the model
declares (``bayes_air/model.py:62-64,71-73,80-82,94-97,106-109``).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch

from .function import simulate_log_joint
from .plan import SimulationPlan


def n_latent(n_airports: int, include_cancellations: bool = False) -> int:
    """``train.py:228-235``."""
    n = 2 * n_airports + n_airports * n_airports
    return n + 2 * n_airports if include_cancellations else n


def map_to_latents(sample: torch.Tensor, n_airports: int, include_cancellations: bool = False
                   ) -> Dict[str, torch.Tensor]:
    """``[P, n_latent]`` log-latents -> constrained latents (tensor form of
    ``map_to_sample_sites``, ``train.py:244-297``)."""
    N = n_airports
    if sample.dim() == 1:
        sample = sample.unsqueeze(0)
    if sample.shape[-1] != n_latent(N, include_cancellations):
        raise ValueError("sample has the wrong number of latent variables")
    out = {"mean_turnaround": torch.exp(sample[:, :N]),
           "mean_service": torch.exp(sample[:, N:2 * N])}
    off = 2 * N
    if include_cancellations:
        out["log_initial_aircraft"] = sample[:, 2 * N:3 * N]
        out["base_cancel_logprob"] = sample[:, 3 * N:4 * N]
        off = 4 * N
    out["travel"] = torch.exp(sample[:, off:].reshape(-1, N, N))
    return out


def _gamma_lp(x, a, b):
    return a * math.log(b) + (a - 1.0) * torch.log(x) - b * x - math.lgamma(a)


def prior_log_prob(lat: Dict[str, torch.Tensor], include_cancellations: bool = False) -> torch.Tensor:
    """Sum of the global latent sites' prior log-probs, ``[P]`` (model.py:58-113); the
    diagonal of the travel-time matrix has no site (``if origin != destination``)."""
    N = lat["mean_service"].shape[1]
    lp = _gamma_lp(lat["mean_turnaround"], 1.0, 2.0).sum(1)
    lp = lp + _gamma_lp(lat["mean_service"], 1.5, 10.0).sum(1)
    off = ~torch.eye(N, dtype=torch.bool, device=lat["travel"].device)
    lp = lp + (_gamma_lp(lat["travel"], 4.0, 1.25) * off).sum((1, 2))
    if include_cancellations:
        c = -0.5 * math.log(2.0 * math.pi)
        lp = lp + (c - 0.5 * lat["log_initial_aircraft"] ** 2).sum(1)
        lp = lp + (c - 0.5 * (lat["base_cancel_logprob"] + 3.0) ** 2).sum(1)
    return lp


class MeanFieldGuide(torch.nn.Module):
    """Diagonal-Normal guide over the log-latent vector.  Surface used by the reference's
    trainer: ``rsample_and_log_prob``, ``sample``, ``log_prob``, ``parameters()``
    (``scripts/training.py:184-294``; zuko's flows are not installable here)."""

    def __init__(self, n: int, init_loc: Optional[torch.Tensor] = None, init_scale: float = 0.1):
        super().__init__()
        self.loc = torch.nn.Parameter(torch.zeros(n) if init_loc is None else init_loc.clone())
        self.log_scale = torch.nn.Parameter(torch.full((n,), math.log(init_scale)))

    def rsample_and_log_prob(self, shape=()) -> Tuple[torch.Tensor, torch.Tensor]:
        shape = (shape,) if isinstance(shape, int) else tuple(shape)
        eps = torch.randn(shape + self.loc.shape, device=self.loc.device, dtype=self.loc.dtype)
        z = self.loc + torch.exp(self.log_scale) * eps
        lp = (-0.5 * eps ** 2 - self.log_scale - 0.5 * math.log(2.0 * math.pi)).sum(-1)
        return z, lp

    def sample(self, shape=()):
        with torch.no_grad():
            return self.rsample_and_log_prob(shape)[0]

    def log_prob(self, z):
        eps = (z - self.loc) * torch.exp(-self.log_scale)
        return (-0.5 * eps ** 2 - self.log_scale - 0.5 * math.log(2.0 * math.pi)).sum(-1)


def model_log_joint(plan: SimulationPlan, sample: torch.Tensor, scales: torch.Tensor,
                    noise: Optional[torch.Tensor] = None, seed: int = 0,
                    check_status: bool = True,
                    day_weights: Optional[torch.Tensor] = None,
                    status_out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``[P]``: what ``trace(condition(model, data)).get_trace(...).log_prob_sum()`` gives
    the reference per particle (``train.py:306-314``), for all P at once.  ``day_weights
    [D]``: multiplicity of each day of the plan in the list of states the reference would be
    given (a bootstrap subsample with repeats, ``training.py:134-136,200-214``, is a weighting
    of the per-day log-probs: days are independent given the latents, ``model.py:126-155``)."""
    lat = map_to_latents(sample, plan.n_airports, plan.include_cancellations)
    lat_airport, lat_route = plan.pack_latents(
        lat["mean_turnaround"], lat["mean_service"], lat["travel"],
        lat.get("log_initial_aircraft"), lat.get("base_cancel_logprob"))
    logp = simulate_log_joint(plan, lat_airport, lat_route, scales, noise, seed=seed,
                              check_status=check_status, status_out=status_out)
    if day_weights is not None:
        logp = logp * day_weights.to(logp)
    return prior_log_prob(lat, plan.include_cancellations) + logp.sum(1)


def batched_elbo(guide, plan: SimulationPlan, scales: torch.Tensor, n_particles: int,
                 seed: Optional[int] = None, check_status: bool = True) -> torch.Tensor:
    """Per-particle ELBO ``[P]`` (``single_particle_elbo``, ``train.py:301-316``)."""
    z, logq = guide.rsample_and_log_prob((n_particles,))
    if seed is None:
        seed = int(torch.randint(0, 2**62, (1,), dtype=torch.int64))
    return model_log_joint(plan, z, scales, None, seed, check_status) - logq


def objective_fn(guide, plan: SimulationPlan, scales: torch.Tensor, n_particles: int,
                 seed: Optional[int] = None) -> torch.Tensor:
    """The reference's loss: ``-mean ELBO / number of flights`` (``train.py:318-326``)."""
    return -batched_elbo(guide, plan, scales, n_particles, seed).mean() / plan.n_flights_total


def posterior_predictive(plan: SimulationPlan, sample: torch.Tensor, scales: torch.Tensor,
                         seed: Optional[int] = None, check_status: bool = True) -> torch.Tensor:
    """Forward-simulates the network for every row of ``sample`` (``[P, n_latent]``
    log-latents, e.g. draws from a trained guide): ``[P, D, Fmax, 3]`` = (cancelled, simulated
    departure time, simulated arrival time).  Flights that carry observations keep them
    (teacher forcing, like the reference with ``obs=`` set); flights without are sampled --
    BASELINE config 4 (delay / cancellation distributions)."""
    lat = map_to_latents(sample, plan.n_airports, plan.include_cancellations)
    lat_airport, lat_route = plan.pack_latents(
        lat["mean_turnaround"], lat["mean_service"], lat["travel"],
        lat.get("log_initial_aircraft"), lat.get("base_cancel_logprob"))
    if seed is None:
        seed = int(torch.randint(0, 2**62, (1,), dtype=torch.int64))
    out = plan.predict(lat_airport.detach(), lat_route.detach(), scales.detach(), None, None, seed=seed)
    if check_status:
        plan.check_status(out["status"])
    return out["sim"]
