"""``torch.autograd.Function`` over the fused CUDA kernels.

forward  -> ``ba_forward``  : simulation + log-joint of every (particle, day), and --
            when any latent requires grad -- the per-(particle, day) gradient rows;
backward -> ``ba_backward`` : contracts those rows with the incoming cotangent.
Replaces the ~2.7 M-node autograd graph the reference builds per particle-day
(``scripts/training.py:289``).
"""
from __future__ import annotations

from typing import Optional

import torch

from .plan import SimulationPlan


class _SimulateLogJoint(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan: SimulationPlan, lat_airport, lat_route, scales, noise, seed,
                noise_is_value, check_status):
        need = any(ctx.needs_input_grad[1:4])
        # per-flight site values given by the caller (value mode) and differentiable: the
        # route of a guide over the per-flight sites (wn_network.py:285-302)
        need_values = bool(noise_is_value and noise is not None and ctx.needs_input_grad[4])
        out = plan.forward(lat_airport.detach(), lat_route.detach(), scales.detach(),
                           None if noise is None else noise.detach(), seed=seed,
                           noise_is_value=noise_is_value, want_grad=need,
                           want_value_grad=need_values)
        if check_status:
            plan.check_status(out["status"])
        ctx.plan = plan
        ctx.tape = out["tape"]
        ctx.value_tape = out["value_tape"]
        ctx.mark_non_differentiable(out["status"])
        return out["logp"], out["status"]

    @staticmethod
    def backward(ctx, dlogp, _dstatus):
        gn = None
        if ctx.value_tape is not None:
            gn = ctx.plan.backward_values(dlogp.contiguous(), ctx.value_tape)
        if ctx.tape is None:
            return (None, None, None, None, gn, None, None, None)
        ga, gr, gs = ctx.plan.backward(dlogp.contiguous(), ctx.tape)
        need = ctx.needs_input_grad
        return (None, ga if need[1] else None, gr if need[2] else None,
                gs.sum(0) if need[3] else None, gn, None, None, None)


def simulate_log_joint(plan: SimulationPlan, lat_airport: torch.Tensor, lat_route: torch.Tensor,
                       scales: torch.Tensor, noise: Optional[torch.Tensor] = None, *,
                       seed: int = 0, noise_is_value: bool = False, check_status: bool = True,
                       status_out: Optional[torch.Tensor] = None):
    """``logp [P, D]``: per-day sum of the 7F per-flight site log-probs.

    ``lat_airport [P,C,N]``, ``lat_route [P,R]``, ``scales [3]`` (all differentiable),
    ``noise [P,D,Fmax,4]`` or None (in-kernel Philox from ``seed``).
    ``check_status=True`` synchronises once to raise on failed particle-days;
    ``status_out`` (a 0-d int32 device tensor) instead accumulates a failure flag on the
    device, to be read when the caller chooses (no host sync per call)."""
    logp, status = _SimulateLogJoint.apply(plan, lat_airport, lat_route, scales, noise, int(seed),
                                      bool(noise_is_value), bool(check_status))
    if status_out is not None:
        status_out.bitwise_or_(status.amax().to(status_out.dtype))
    return logp
