"""``pyro.infer`` / ``pyro.optim`` surface of the mini runtime: what the reference's second
call site uses (``scripts/wn/wn_network.py:285-302``):

    model = pyro.poutine.scale(air_traffic_network_model, scale=1.0 / days)
    guide = pyro.infer.autoguide.AutoMultivariateNormal(model)
    svi = pyro.infer.SVI(model, guide, pyro.optim.ClippedAdam({"lr": lr, "lrd": lrd}),
                         pyro.infer.Trace_ELBO())
    loss = svi.step(states, dt)

Semantics follow Pyro's documented behaviour (re-implemented, not copied): the autoguides
discover the model's latent sample sites by tracing it once, keep one Normal / multivariate
Normal over the concatenated UNCONSTRAINED site values (``biject_to(support)``), and replay
constrained values into the model; ``Trace_ELBO`` is ``E_q[log p(x, z) - log q(z)]`` with
every site's ``scale`` applied (``poutine.scale``), estimated with ``num_particles`` draws
-- as a Python loop or, with ``vectorize_particles=True``, as ONE model call whose site
values carry a leading particle dimension (the drop-in model batches over it).
"""
from __future__ import annotations

import math
import types
from collections import OrderedDict
from typing import Callable, Dict, Optional

import torch
from torch.distributions import biject_to, constraints

from . import _mini as pyro


# ---------------------------------------------------------------------------------- optim
class _Optim:
    """Pyro-style optimiser wrapper: built from a dict of arguments, owns one torch optimiser
    over whatever parameters the param store holds when ``__call__`` first sees them."""

    def __init__(self, optim_args: Dict, clip_norm: Optional[float] = None, lrd: float = 1.0):
        self.args = {k: v for k, v in optim_args.items() if k not in ("clip_norm", "lrd")}
        self.clip_norm = optim_args.get("clip_norm", clip_norm)
        self.lrd = optim_args.get("lrd", lrd)
        self._opt: Optional[torch.optim.Optimizer] = None
        self._known: Dict[int, torch.Tensor] = {}

    def __call__(self, params):
        params = [p for p in params if p.grad is not None or id(p) in self._known]
        fresh = [p for p in params if id(p) not in self._known]
        for p in fresh:
            self._known[id(p)] = p
        if self._opt is None:
            self._opt = torch.optim.Adam(fresh, **self.args)
        elif fresh:
            self._opt.add_param_group({"params": fresh})
        if self.clip_norm is not None:
            for p in self._known.values():          # ClippedAdam clips per coordinate
                if p.grad is not None:
                    p.grad.clamp_(-self.clip_norm, self.clip_norm)
        self._opt.step()
        if self.lrd != 1.0:
            for g in self._opt.param_groups:
                g["lr"] *= self.lrd

    def get_state(self):
        return None if self._opt is None else self._opt.state_dict()


def Adam(optim_args):  # noqa: N802
    return _Optim(optim_args)


def ClippedAdam(optim_args):  # noqa: N802
    return _Optim(optim_args, clip_norm=10.0, lrd=1.0)


optim = types.SimpleNamespace(Adam=Adam, ClippedAdam=ClippedAdam)


# ------------------------------------------------------------------------------ autoguides
class _AutoGuide:
    """Base of the autoguides: latent-site discovery, (un)constraining, replay."""

    def __init__(self, model: Callable, init_scale: float = 0.1, prefix: str = "auto"):
        self.model, self.init_scale, self.prefix = model, float(init_scale), prefix
        self._sites: "OrderedDict[str, dict]" = OrderedDict()
        self.latent_dim = 0

    def _setup(self, *args, **kwargs):
        with pyro.poutine.block():
            tr = pyro.poutine.trace(self.model).get_trace(*args, **kwargs)
        off = 0
        for name, site in tr.iter_stochastic_nodes():
            fn = site["fn"]
            support = getattr(fn, "support", constraints.real)
            t = biject_to(support)
            with torch.no_grad():
                draws = torch.stack([fn() for _ in range(33)]) if hasattr(fn, "batch_shape") else \
                    site["value"].detach().unsqueeze(0)
                med = draws.median(0).values          # init_to_median
                unc = t.inv(med)
            n = unc.numel()
            self._sites[name] = {"transform": t, "shape": unc.shape, "slice": slice(off, off + n),
                                 "init": unc.reshape(-1)}
            off += n
        if off == 0:
            raise RuntimeError("the model has no latent sample sites to guide")
        self.latent_dim = off
        self._init_loc = torch.cat([s["init"] for s in self._sites.values()])

    def _latent(self, sample_shape):                   # -> (z [.., latent_dim], log q [..])
        raise NotImplementedError

    def __call__(self, *args, **kwargs):
        if not self._sites:
            self._setup(*args, **kwargs)
        shape = tuple(getattr(pyro, "_PARTICLE_SHAPE", ()))
        z, logq = self._latent(shape)
        pyro.sample(f"_{self.prefix}_latent", pyro._Delta(z, log_density=logq, event_dim=1), infer={"is_auxiliary": True})
        out = {}
        for name, s in self._sites.items():
            zi = z[..., s["slice"]].reshape(shape + tuple(s["shape"]))
            v = s["transform"](zi)
            ladj = s["transform"].log_abs_det_jacobian(zi, v)
            ladj = ladj.reshape(shape + (-1,)).sum(-1) if ladj.dim() > len(shape) else ladj
            out[name] = pyro.sample(name, pyro._Delta(v, log_density=-ladj, event_dim=len(s["shape"])))
        return out

    def median(self, *args, **kwargs):
        if not self._sites:
            self._setup(*args, **kwargs)
        loc = self._loc().detach()
        return {n: s["transform"](loc[s["slice"]].reshape(s["shape"])) for n, s in self._sites.items()}


class AutoNormal(_AutoGuide):
    """Diagonal Normal over the unconstrained latents."""

    def _loc(self):
        return pyro.param(f"{self.prefix}_loc", lambda: self._init_loc.clone())

    def _latent(self, shape):
        loc = self._loc()
        scale = pyro.param(f"{self.prefix}_scale", lambda: torch.full_like(self._init_loc, self.init_scale),
                           constraint=constraints.positive)
        eps = torch.randn(shape + loc.shape, device=loc.device, dtype=loc.dtype)
        z = loc + scale * eps
        logq = (-0.5 * eps * eps - torch.log(scale) - 0.5 * math.log(2.0 * math.pi)).sum(-1)
        return z, logq


AutoDiagonalNormal = AutoNormal


class AutoMultivariateNormal(_AutoGuide):
    """Full-covariance Normal (Cholesky factor ``scale_tril``) over the unconstrained
    latents.  At the WN network size the factor has (2N + N^2)^2 / 2 entries: the reference
    uses this guide on top-k sub-networks only (``wn_network.py:222-238``)."""

    def _loc(self):
        return pyro.param(f"{self.prefix}_loc", lambda: self._init_loc.clone())

    def _latent(self, shape):
        loc = self._loc()
        tril = pyro.param(f"{self.prefix}_scale_tril",
                          lambda: torch.eye(self.latent_dim, device=self._init_loc.device) * self.init_scale,
                          constraint=constraints.lower_cholesky)
        eps = torch.randn(shape + loc.shape, device=loc.device, dtype=loc.dtype)
        z = loc + eps @ tril.T
        logq = (-0.5 * eps * eps).sum(-1) - torch.log(torch.diagonal(tril)).sum() \
            - 0.5 * self.latent_dim * math.log(2.0 * math.pi)
        return z, logq


class AutoDelta(_AutoGuide):
    """Point estimate (MAP) of every latent site."""

    def _loc(self):
        return pyro.param(f"{self.prefix}_loc", lambda: self._init_loc.clone())

    def _latent(self, shape):
        loc = self._loc()
        z = loc.expand(shape + loc.shape)
        return z, torch.zeros(shape, device=loc.device, dtype=loc.dtype)


autoguide = types.SimpleNamespace(AutoNormal=AutoNormal, AutoDiagonalNormal=AutoDiagonalNormal,
                                  AutoMultivariateNormal=AutoMultivariateNormal, AutoDelta=AutoDelta)


# ------------------------------------------------------------------------------------ ELBO
class Trace_ELBO:  # noqa: N801
    def __init__(self, num_particles: int = 1, vectorize_particles: bool = False,
                 max_plate_nesting=None, **_unused):
        self.num_particles, self.vectorize = int(num_particles), bool(vectorize_particles)

    def _one(self, model, guide, args, kwargs):
        guide_tr = pyro.poutine.trace(guide).get_trace(*args, **kwargs)
        model_tr = pyro.poutine.trace(pyro.poutine.replay(model, trace=guide_tr)).get_trace(*args, **kwargs)
        return model_tr.log_prob_sum() - guide_tr.log_prob_sum()

    def differentiable_loss(self, model, guide, *args, **kwargs):
        if self.vectorize and self.num_particles > 1:
            pyro._PARTICLE_SHAPE = (self.num_particles,)
            try:
                elbo = self._one(model, guide, args, kwargs) / self.num_particles
            finally:
                pyro._PARTICLE_SHAPE = ()
            return -elbo
        elbo = 0.0
        for _ in range(self.num_particles):
            elbo = elbo + self._one(model, guide, args, kwargs) / self.num_particles
        return -elbo

    def loss(self, model, guide, *args, **kwargs):
        with torch.no_grad():
            return float(self.differentiable_loss(model, guide, *args, **kwargs))

    def loss_and_grads(self, model, guide, *args, **kwargs):
        loss = self.differentiable_loss(model, guide, *args, **kwargs)
        loss.backward()
        return float(loss.detach())


class SVI:
    def __init__(self, model, guide, optim, loss, **_unused):  # noqa: A002
        self.model, self.guide, self.optim, self.loss = model, guide, optim, loss

    def step(self, *args, **kwargs) -> float:
        store = pyro.get_param_store()
        for _, raw in store.named_parameters():
            raw.grad = None
        loss = self.loss.loss_and_grads(self.model, self.guide, *args, **kwargs)
        self.optim([raw for _, raw in store.named_parameters()])
        return loss

    def evaluate_loss(self, *args, **kwargs) -> float:
        return self.loss.loss(self.model, self.guide, *args, **kwargs)


infer = types.SimpleNamespace(SVI=SVI, Trace_ELBO=Trace_ELBO, autoguide=autoguide)
