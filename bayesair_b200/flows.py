"""Conditional normalizing-flow guide for the CalNF objective.

The reference trains ``zuko.flows.NSF(features=2N+N^2, context=K, hidden_features=(64, 64))``
(``scripts/wn/train.py:622-626``) and only ever touches this surface of it
(``scripts/training.py:184-294``, ``scripts/utils.py:16-26``):

    dist = guide(label)                       # label: [K] mixture / permutation label
    z, logq = dist.rsample_and_log_prob((P,)) # reparameterised draws + their log-density
    dist.log_prob(z); dist.sample((P,)); guide.parameters()

zuko is not installable here and its NSF is autoregressive (one conditioner pass per feature
when sampling: 10 200 sequential passes at the WN network size), so this is a from-scratch
flow with the same surface and the same building block -- monotonic rational-quadratic
splines (Durkan et al. 2019) with ``bins`` knots on ``[-bound, bound]`` and identity tails
-- arranged as COUPLING layers: each layer transforms one half of the features with spline
parameters produced by an MLP of the other half and the context, so sampling and density
evaluation are both one pass.  Sampling runs the splines forward (no root finding on the
training path); ``log_prob`` of a given point uses the analytic inverse.

The spline itself (the only non-GEMM work: ``3 bins - 1`` parameters per feature and sample)
runs as ONE fused kernel pair of the library when the tensors are on a GPU
(``ba_rqs_forward`` / ``ba_rqs_backward``, csrc/ba_guide.cu); the pure-torch restatement
below is what the CPU tests check it against.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

MIN_BIN = 1e-3          # smallest relative bin width / height
MIN_SLOPE = 1e-3        # smallest knot derivative


def _knots(params: torch.Tensor, bins: int, bound: float):
    """Unconstrained ``[..., 3 bins - 1]`` -> knot positions ``xk, yk [..., bins + 1]`` and
    knot derivatives ``dk [..., bins + 1]`` (1 at both ends: identity tails)."""
    w, h, d = params[..., :bins], params[..., bins:2 * bins], params[..., 2 * bins:]
    w = MIN_BIN + (1.0 - MIN_BIN * bins) * torch.softmax(w, -1)
    h = MIN_BIN + (1.0 - MIN_BIN * bins) * torch.softmax(h, -1)
    zero = torch.zeros_like(w[..., :1])
    xk = torch.cat([zero, torch.cumsum(w, -1)], -1) * (2.0 * bound) - bound
    yk = torch.cat([zero, torch.cumsum(h, -1)], -1) * (2.0 * bound) - bound
    xk = torch.cat([xk[..., :-1], torch.full_like(zero, bound)], -1)
    yk = torch.cat([yk[..., :-1], torch.full_like(zero, bound)], -1)
    # softplus(d + c) with softplus(c) = 1 - MIN_SLOPE: zero parameters give the identity
    c = math.log(math.expm1(1.0 - MIN_SLOPE))
    dk = MIN_SLOPE + F.softplus(d + c)
    one = torch.ones_like(zero)
    dk = torch.cat([one, dk, one], -1)
    return xk, yk, dk


def rqs_torch(x: torch.Tensor, params: torch.Tensor, bins: int, bound: float,
              inverse: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """Monotonic rational-quadratic spline, elementwise: ``x [...]``, ``params [..., 3 bins -
    1]`` -> ``(y, log |dy/dx|)``; with ``inverse`` the map runs backwards and the second
    output is ``log |dx/dy|`` of the point returned."""
    xk, yk, dk = _knots(params, bins, bound)
    inside = (x > -bound) & (x < bound)
    xc = x.clamp(-bound, bound)
    grid = yk if inverse else xk
    k = (torch.searchsorted(grid.contiguous(), xc.unsqueeze(-1).contiguous(), right=True) - 1)
    k = k.clamp(0, bins - 1)
    g = lambda t, i=0: torch.gather(t, -1, k + i).squeeze(-1)  # noqa: E731
    x0, x1, y0, y1, d0, d1 = g(xk), g(xk, 1), g(yk), g(yk, 1), g(dk), g(dk, 1)
    wk, hk = x1 - x0, y1 - y0
    s = hk / wk
    if not inverse:
        th = (xc - x0) / wk
        t1 = th * (1.0 - th)
        den = s + (d1 + d0 - 2.0 * s) * t1
        y = y0 + hk * (s * th * th + d0 * t1) / den
        num = s * s * (d1 * th * th + 2.0 * s * t1 + d0 * (1.0 - th) ** 2)
        ld = torch.log(num) - 2.0 * torch.log(den)
    else:
        dy = xc - y0
        e = d1 + d0 - 2.0 * s
        a = hk * (s - d0) + dy * e
        b = hk * d0 - dy * e
        c = -s * dy
        disc = (b * b - 4.0 * a * c).clamp_min(0.0)
        th = 2.0 * c / (-b - torch.sqrt(disc))
        t1 = th * (1.0 - th)
        den = s + e * t1
        y = x0 + th * wk
        num = s * s * (d1 * th * th + 2.0 * s * t1 + d0 * (1.0 - th) ** 2)
        ld = -(torch.log(num) - 2.0 * torch.log(den))
    y = torch.where(inside, y, x)
    ld = torch.where(inside, ld, torch.zeros_like(ld))
    return y, ld


class _RqsCuda(torch.autograd.Function):
    """The fused library kernels (forward map only; the inverse is never differentiated on
    the training path and stays in torch)."""

    @staticmethod
    def forward(ctx, x, params, bins, bound):
        from . import _capi
        x = x.contiguous()
        params = params.contiguous()
        y = torch.empty_like(x)
        ld = torch.empty_like(x)
        _capi.rqs_forward(x, params, y, ld, bins, bound)
        ctx.save_for_backward(x, params)
        ctx.bins, ctx.bound = bins, bound
        return y, ld

    @staticmethod
    def backward(ctx, gy, gld):
        from . import _capi
        x, params = ctx.saved_tensors
        gx = torch.empty_like(x)
        gp = torch.empty_like(params)
        _capi.rqs_backward(x, params, gy.contiguous(), gld.contiguous(), gx, gp, ctx.bins, ctx.bound)
        return gx, gp, None, None


def rqs(x, params, bins, bound, inverse=False):
    if x.is_cuda and not inverse and x.dtype == torch.float32:
        return _RqsCuda.apply(x, params, bins, float(bound))
    return rqs_torch(x, params, bins, bound, inverse)


class _Coupling(torch.nn.Module):
    def __init__(self, idx_a: torch.Tensor, idx_b: torch.Tensor, context: int,
                 hidden: Sequence[int], bins: int):
        super().__init__()
        self.register_buffer("idx_a", idx_a)
        self.register_buffer("idx_b", idx_b)
        self.register_buffer("inv", torch.argsort(torch.cat([idx_a, idx_b])))
        dims = [idx_a.numel() + context] + list(hidden)
        self.hidden = torch.nn.ModuleList(torch.nn.Linear(i, o) for i, o in zip(dims[:-1], dims[1:]))
        self.out = torch.nn.Linear(dims[-1], idx_b.numel() * (3 * bins - 1))
        torch.nn.init.zeros_(self.out.weight)        # the flow starts as the identity
        torch.nn.init.zeros_(self.out.bias)
        self.bins = bins

    def spline_params(self, xa: torch.Tensor, context: torch.Tensor) -> torch.Tensor:
        h = torch.cat([xa, context.expand(xa.shape[:-1] + context.shape[-1:])], -1)
        for lin in self.hidden:
            h = F.elu(lin(h))
        return self.out(h).reshape(xa.shape[:-1] + (self.idx_b.numel(), 3 * self.bins - 1))

    def merge(self, xa: torch.Tensor, xb: torch.Tensor) -> torch.Tensor:
        return torch.cat([xa, xb], -1)[..., self.inv]


class FlowDistribution:
    """``guide(label)``: the flow at one context.  ``rsample_and_log_prob``, ``rsample``,
    ``sample``, ``log_prob`` -- the calls of ``scripts/training.py`` / ``scripts/utils.py``."""

    def __init__(self, flow: "ConditionalSplineFlow", context: torch.Tensor):
        self.flow, self.context = flow, context

    def rsample_and_log_prob(self, shape=()) -> Tuple[torch.Tensor, torch.Tensor]:
        shape = (shape,) if isinstance(shape, int) else tuple(shape)
        fl = self.flow
        u = torch.randn(shape + (fl.features,), device=fl.loc.device, dtype=fl.loc.dtype)
        logq = (-0.5 * u * u).sum(-1) - 0.5 * fl.features * math.log(2.0 * math.pi)
        for layer in fl.layers:
            ua = u[..., layer.idx_a]
            yb, ld = rqs(u[..., layer.idx_b], layer.spline_params(ua, self.context), fl.bins, fl.bound)
            u = layer.merge(ua, yb)
            logq = logq - ld.sum(-1)
        x = fl.loc + fl.scale * u
        return x, logq - fl.log_scale_sum

    def rsample(self, shape=()):
        return self.rsample_and_log_prob(shape)[0]

    def sample(self, shape=()):
        with torch.no_grad():
            return self.rsample_and_log_prob(shape)[0]

    def log_prob(self, x: torch.Tensor) -> torch.Tensor:
        fl = self.flow
        u = (x - fl.loc) / fl.scale
        logq = -fl.log_scale_sum
        for layer in reversed(fl.layers):
            ua = u[..., layer.idx_a]
            zb, ld = rqs(u[..., layer.idx_b], layer.spline_params(ua, self.context), fl.bins,
                         fl.bound, inverse=True)
            u = layer.merge(ua, zb)
            logq = logq + ld.sum(-1)
        return logq + (-0.5 * u * u).sum(-1) - 0.5 * fl.features * math.log(2.0 * math.pi)


class ConditionalSplineFlow(torch.nn.Module):
    """``features``-dimensional flow conditioned on a ``context``-dimensional label.

    ``transforms`` coupling layers with alternating halves (even / odd feature indices, then
    first / second half, so every feature conditions on every other after four layers),
    conditioner MLPs with ``hidden_features`` (ELU), ``bins`` spline knots on
    ``[-bound, bound]``.  ``init_loc`` / ``init_scale`` are a FIXED affine map applied after
    the flow (x = loc + scale u): the log-latents of the air-traffic model live around
    log(0.01) .. log(3), far from a unit Gaussian."""

    def __init__(self, features: int, context: int, hidden_features: Sequence[int] = (64, 64),
                 transforms: int = 4, bins: int = 8, bound: float = 5.0,
                 init_loc: Optional[torch.Tensor] = None, init_scale=1.0):
        super().__init__()
        if features < 2:
            raise ValueError("a coupling flow needs at least two features")
        self.features, self.context_dim, self.bins, self.bound = features, context, bins, float(bound)
        idx = torch.arange(features)
        halves = [(idx[0::2], idx[1::2]), (idx[1::2], idx[0::2]),
                  (idx[:features // 2], idx[features // 2:]), (idx[features // 2:], idx[:features // 2])]
        self.layers = torch.nn.ModuleList(
            _Coupling(*halves[i % 4], context, hidden_features, bins) for i in range(transforms))
        loc = torch.zeros(features) if init_loc is None else torch.as_tensor(init_loc, dtype=torch.float32)
        scale = torch.as_tensor(init_scale, dtype=torch.float32).expand(features).clone()
        self.register_buffer("loc", loc.clone())
        self.register_buffer("scale", scale)

    @property
    def log_scale_sum(self):
        return torch.log(self.scale).sum()

    def forward(self, context: torch.Tensor) -> FlowDistribution:
        context = torch.as_tensor(context, dtype=self.loc.dtype, device=self.loc.device)
        if context.shape[-1] != self.context_dim:
            raise ValueError("context has the wrong length")
        return FlowDistribution(self, context)
