"""Drop-in ``air_traffic_network_model`` backed by the sm_100a kernels.

Same signature, same ``pyro.param`` names/constraints, same global latent
``pyro.sample`` site names and priors as the reference model
(``bayes_air/model.py:14-121``), so code that conditions on / guides those sites
(``map_to_sample_sites`` in ``scripts/wn/train.py:244-297``, autoguides in
``scripts/wn/wn_network.py:285-302``) runs unchanged.  Everything per flight --
the 7 sample sites per flight of ``bayes_air/network.py:102-109,158-163`` and
``bayes_air/types/airport.py:144-220`` and the tick loop around them
(``model.py:126-209``) -- is replaced by ONE ``pyro.factor`` whose value is the
sum of those sites' log-probs, computed by ``libbayesair_b200.so``.

Batched use: every global latent may be 0-d (one particle, like the reference)
or ``[P]`` (``map_to_sample_sites`` already produces ``[P]`` values for a
``[P, n_latent]`` guide sample); the factor then has shape ``[P]``.
"""
from __future__ import annotations

import contextlib
import threading
from typing import List, Optional

import torch

from .function import simulate_log_joint
from .plan import plan_for_states
from .ppl import dist, pyro

FAR_FUTURE_TIME = 30.0     # bayes_air/model.py:11
FACTOR_NAME = "air_traffic_log_joint"

_tls = threading.local()


@contextlib.contextmanager
def flight_noise(noise: Optional[torch.Tensor] = None, *, seed: Optional[int] = None,
                 noise_is_value: bool = False):
    """Pins the per-flight randomness of model calls made inside the block.

    ``noise``: ``[P, D, Fmax, 4]`` standard draws (Exp(1), N(0,1), Exp(1), N(0,1)) per
    flight in sorted pending order -- the eps behind the reference's four latent
    per-flight sites (``*_departure_service_time``, ``*_travel_time``,
    ``*_arrival_service_time``, ``*_turnaround_time``).  With ``noise_is_value=True``
    the tensor holds the site values themselves.  ``seed``: Philox seed when no
    tensor is given.  Outside such a block the seed comes from the torch RNG."""
    prev = getattr(_tls, "override", None)
    _tls.override = (noise, seed, noise_is_value)
    try:
        yield
    finally:
        _tls.override = prev


def _as_batch(v: torch.Tensor, device) -> torch.Tensor:
    v = torch.as_tensor(v)
    if v.device != device or v.dtype != torch.float32:
        v = v.to(device=device, dtype=torch.float32)
    return v.reshape(-1)


def _stack_sites(values: List[torch.Tensor], device):
    vals = [_as_batch(v, device) for v in values]
    P = max(v.shape[0] for v in vals)
    vals = [v if v.shape[0] == P else v.expand(P) for v in vals]
    return torch.stack(vals, dim=1), P     # [P, n]


def _flight_sites(plan, rows, travel, codes, v_t, v_u, P, batched, device):
    """Declares the four latent per-flight sites of every non-cancelled flight and returns
    their values as the kernel's ``[P, D, Fmax, 4]`` value tensor plus the sum of their
    log-probs ``[P]`` (what the kernel's total contains besides the observed sites)."""
    turnaround, service = rows[0], rows[1]
    D, F = plan.n_days, plan.max_flights
    cols = [[None] * F for _ in range(D)]
    latent_lp = torch.zeros((P,), device=device)
    one = torch.ones((P,), device=device)
    for d, spec in enumerate(plan.specs):
        for f in range(spec.n_flights):
            if spec.cancel_obs[f] == 1:
                continue
            o, dd = int(spec.airport_global[spec.origin[f]]), int(spec.airport_global[spec.dest[f]])
            stem = f"day{d}_{spec.flight_names[f]}"
            ms_o, ms_d = _as_batch(service[o], device), _as_batch(service[dd], device)
            mt_d = _as_batch(turnaround[dd], device)
            T = _as_batch(travel[(codes[o], codes[dd])], device)
            sites = (
                ("_departure_service_time", dist.Exponential(1.0 / ms_o.expand(P))),
                ("_travel_time", dist.Normal(T.expand(P), T.expand(P) * v_t)),
                ("_arrival_service_time", dist.Exponential(1.0 / ms_d.expand(P))),
                ("_turnaround_time", dist.Normal(mt_d.expand(P), v_u * mt_d.expand(P))),
            )
            vals = []
            for suffix, fn in sites:
                v = _as_batch(pyro.sample(stem + suffix, fn), device).expand(P)
                latent_lp = latent_lp + fn.log_prob(v)
                vals.append(v)
            cols[d][f] = torch.stack(vals, dim=-1)                 # [P, 4]
    pad = torch.stack([one, one, one, one], dim=-1)
    values = torch.stack([torch.stack([c if c is not None else pad for c in day], dim=1)
                          for day in cols], dim=1)                 # [P, D, F, 4]
    return values.contiguous(), latent_lp


def air_traffic_network_model(
    states,
    delta_t: float = 0.1,
    max_t: float = FAR_FUTURE_TIME,
    device=None,
    include_cancellations: bool = False,
    expose_flight_sites: bool = False,
):
    """See the module docstring.  ``device`` must be a CUDA device (``None`` ->
    the current one); the simulation has no CPU path.  Returns ``states``
    unmodified (every caller of the reference ignores the return value); if some
    flight has no observations (``obs=None``) the call is a posterior-predictive
    simulation instead and returns ``[P, D, Fmax, 3]`` (cancelled, departure time,
    arrival time) -- the reference returns the mutated state objects there.

    ``expose_flight_sites=True`` (an extension; plans without cancellations, every flight
    observed) declares the four LATENT per-flight sites of the reference as well --
    ``day{d}_{flight}_departure_service_time``, ``_travel_time``, ``_arrival_service_time``,
    ``_turnaround_time`` with the reference's distributions (``types/airport.py:144-147,
    217-220``, ``network.py:158-163``) -- so that a guide over EVERY latent site
    (``AutoMultivariateNormal(model)``, ``scripts/wn/wn_network.py:285-302``) proposes them;
    the kernel then runs in value mode on the proposed values, the factor carries only the
    observed sites, and gradients reach the values through the kernel's value tape.  This is
    4F Python statements per day: for the small networks that route is viable for."""
    plan = plan_for_states(states, delta_t, max_t, include_cancellations, device)
    device = plan.device

    def scalar(x):
        return torch.tensor(x, device=device)

    # system-level parameters (model.py:41-55)
    sigma = pyro.param("runway_use_time_std_dev", scalar(0.1), constraint=dist.constraints.positive)
    v_t = pyro.param("travel_time_variation", scalar(0.1), constraint=dist.constraints.positive)
    v_u = pyro.param("turnaround_time_variation", scalar(0.1), constraint=dist.constraints.positive)

    codes = plan.global_codes
    N = len(codes)
    # global latent sites, same names / priors / order as model.py:58-113
    turn_prior = dist.Gamma(scalar(1.0), scalar(2.0))
    serv_prior = dist.Gamma(scalar(1.5), scalar(10.0))
    trav_prior = dist.Gamma(scalar(4.0), scalar(1.25))
    turnaround = [pyro.sample(f"{c}_mean_turnaround_time", turn_prior) for c in codes]
    service = [pyro.sample(f"{c}_mean_service_time", serv_prior) for c in codes]
    travel = {}
    for o in codes:
        for d in codes:
            if o != d:
                travel[(o, d)] = pyro.sample(f"travel_time_{o}_{d}", trav_prior)
    rows = [turnaround, service]
    if include_cancellations:
        n0_prior = dist.Normal(scalar(0.0), scalar(1.0))
        bc_prior = dist.Normal(scalar(-3.0), scalar(1.0))
        rows.append([pyro.sample(f"{c}_log_initial_available_aircraft", n0_prior) for c in codes])
        rows.append([pyro.sample(f"{c}_base_cancel_logprob", bc_prior) for c in codes])

    # kernel layout: [P, C, N] airport rows, [P, R] used routes
    stacked = [_stack_sites(r, device) for r in rows]
    ro = plan._h.route_origin
    rd = plan._h.route_dest
    route_vals, Pr = _stack_sites([travel[(codes[o], codes[d])] for o, d in zip(ro, rd)], device) \
        if len(ro) else (torch.zeros((1, 0), device=device), 1)
    P = max([p for _, p in stacked] + [Pr])
    batched = any(torch.as_tensor(v).dim() > 0 for r in rows for v in r) or \
        any(torch.as_tensor(v).dim() > 0 for v in travel.values())
    lat_airport = torch.stack([s if s.shape[0] == P else s.expand(P, N) for s, _ in stacked], dim=1)
    lat_route = route_vals if route_vals.shape[0] == P else route_vals.expand(P, route_vals.shape[1])
    scales = torch.stack([sigma.reshape(()), v_t.reshape(()), v_u.reshape(())]).to(device)

    noise, seed, is_value = getattr(_tls, "override", None) or (None, None, False)
    if noise is None and seed is None:
        seed = int(torch.randint(0, 2**62, (1,), dtype=torch.int64))
    if plan.has_unobserved:
        # obs=None somewhere: posterior-predictive run.  Like the reference, the observed
        # sites of such flights are sampled; there is no log-joint to report.  The outcomes
        # are exposed as a deterministic site and returned.
        out = plan.predict(lat_airport.detach(), lat_route.detach(), scales.detach(), noise, None,
                           seed=seed or 0)
        plan.check_status(out["status"])
        sim = out["sim"] if batched else out["sim"][0]
        pyro.deterministic("air_traffic_simulated", sim)
        return sim
    if expose_flight_sites:
        if include_cancellations or noise is not None:
            raise ValueError("expose_flight_sites: plans without cancellations, no flight_noise override")
        values, latent_lp = _flight_sites(plan, rows, travel, codes, v_t, v_u, P, batched, device)
        logp = simulate_log_joint(plan, lat_airport, lat_route, scales, values, noise_is_value=True)
        # the per-flight latent sites are ordinary sites now: the factor keeps the observed ones
        log_joint = logp.sum(dim=1) - latent_lp
        pyro.factor(FACTOR_NAME, log_joint if batched else log_joint.reshape(()))
        return states
    logp = simulate_log_joint(plan, lat_airport, lat_route, scales, noise,
                              seed=seed or 0, noise_is_value=is_value)
    log_joint = logp.sum(dim=1)               # sum over days -> [P]
    pyro.factor(FACTOR_NAME, log_joint if batched else log_joint.reshape(()))
    return states
