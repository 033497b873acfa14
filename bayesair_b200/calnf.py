"""CalNF training step over the fused kernels (BASELINE configs[2]).

Batched restatement of the reference's trainer loop (``scripts/training.py:177-294``) for the
air-traffic objective (``scripts/wn/train.py:301-326``).  Per optimisation step the reference
evaluates, each as a Python loop over particles of whole-network Pyro traces:

  * ``calibration_substeps`` updates of the mixture label on the failure ELBO
    (``training.py:180-191``);
  * one ELBO per bootstrap permutation ``j`` of the failure days with the one-hot label
    ``e_j`` (``training.py:200-217``; permutations drawn with replacement, ``:134-136``);
  * the nominal ELBO with the zero label (``training.py:249-255``);
  * the calibration term: the failure ELBO at the current mixture label (``:266-270``);
  * gradient-norm clipping, Adam, StepLR (``:291-296``).

Here every ELBO is ONE guide draw of ``n_particles`` log-latent vectors and ONE ``ba_forward``
over (particles x days) of a compiled plan; a permutation with repeats is a weighting of the
per-day log-probs.  Across GPUs the particles are sharded and the only exchange is one sum
all-reduce of the flat gradient buffer the guide's parameters live in (``FlatParameters``):
with a flow guide that buffer is tens of MB, the bandwidth-bound collective of this path.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist

from .objective import model_log_joint
from .plan import SimulationPlan


class FlatParameters:
    """Re-homes every parameter of ``module`` (and its gradient) as a view into ONE flat fp32
    buffer, so that the gradient all-reduce is a single collective on memory autograd
    already wrote (no ``torch.cat``, no per-parameter copies) and the optimiser updates one
    tensor."""

    def __init__(self, module: torch.nn.Module):
        params = [p for p in module.parameters() if p.requires_grad]
        if not params:
            raise ValueError("module has no trainable parameters")
        dev, n = params[0].device, sum(p.numel() for p in params)
        self.flat = torch.empty(n, device=dev, dtype=torch.float32)
        self.grad = torch.zeros(n, device=dev, dtype=torch.float32)
        off = 0
        with torch.no_grad():
            for p in params:
                m = p.numel()
                self.flat[off:off + m].copy_(p.reshape(-1))
                p.data = self.flat[off:off + m].view(p.shape)
                p.grad = self.grad[off:off + m].view(p.shape)
                off += m
        self.params = params
        self.parameter = torch.nn.Parameter(self.flat, requires_grad=True)
        self.parameter.grad = self.grad

    def zero_grad(self):
        self.grad.zero_()              # in place: the views stay attached

    def all_reduce(self, group=None):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            with torch.cuda.nvtx.range("ba.calnf.allreduce") if self.grad.is_cuda else _null():
                dist.all_reduce(self.grad, group=group)

    def clip_grad_norm_(self, max_norm: float) -> torch.Tensor:
        norm = torch.linalg.vector_norm(self.grad)
        self.grad.mul_(torch.clamp(max_norm / (norm + 1e-6), max=1.0))
        return norm

    @property
    def nbytes(self) -> int:
        return self.grad.numel() * 4


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def elbo_loss(guide_dist, plan: SimulationPlan, scales: torch.Tensor, n_particles: int,
              day_weights: Optional[torch.Tensor] = None, seed: Optional[int] = None,
              status: Optional[torch.Tensor] = None, world_size: int = 1) -> torch.Tensor:
    """The reference's ``objective_fn(guide_dist, n, states)`` (``train.py:318-326``):
    ``-mean_p ELBO_p / number of flights`` over this rank's ``n_particles`` (divided by
    ``world_size`` so that the sum over ranks is the global mean)."""
    z, logq = guide_dist.rsample_and_log_prob((n_particles,))
    if seed is None:
        seed = int(torch.randint(0, 2**62, (1,), dtype=torch.int64))
    if day_weights is None:
        flights = float(plan.n_flights_total)
    else:
        per_day = torch.tensor([s.n_flights for s in plan.specs], dtype=torch.float32)
        flights = float((per_day * day_weights.detach().cpu().float()).sum())
    elbo = model_log_joint(plan, z, scales, None, seed, check_status=status is None,
                           day_weights=day_weights, status_out=status) - logq
    return -elbo.mean() / (flights * world_size)


class CalNFTrainer:
    """One object per training run; ``step()`` is one iteration of ``training.py:177-296``.

    ``failure_plan`` / ``nominal_plan``: compiled plans of the failure / nominal days;
    ``guide``: ``label [K] -> distribution`` (``flows.ConditionalSplineFlow``)."""

    def __init__(self, guide: torch.nn.Module, failure_plan: SimulationPlan,
                 nominal_plan: Optional[SimulationPlan], scales: torch.Tensor, *,
                 n_particles: int, n_permutations: int, lr: float = 1e-3, lr_gamma: float = 1.0,
                 lr_steps: int = 1000, grad_clip: float = 10.0, weight_decay: float = 0.0,
                 calibration_weight: float = 1.0, elbo_weight: float = 1.0,
                 calibration_lr: float = 1e-3, calibration_substeps: int = 1,
                 calibrate: bool = True, exclude_nominal: bool = False, seed: int = 0,
                 group=None):
        self.guide, self.fplan, self.nplan = guide, failure_plan, nominal_plan
        self.scales, self.P, self.K = scales, int(n_particles), int(n_permutations)
        self.calibrate, self.exclude_nominal = calibrate, exclude_nominal or nominal_plan is None
        self.calibration_weight, self.elbo_weight = calibration_weight, elbo_weight
        self.grad_clip, self.substeps, self.group = grad_clip, calibration_substeps, group
        dev = failure_plan.device
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        self.fp = FlatParameters(guide)
        self.opt = torch.optim.Adam([self.fp.parameter], lr=lr, weight_decay=weight_decay,
                                    fused=dev.type == "cuda")
        self.sched = torch.optim.lr_scheduler.StepLR(self.opt, step_size=lr_steps, gamma=lr_gamma)
        # training.py:114-126: the mixture label, frozen at uniform when its lr is zero
        self.mixture_label = torch.full((self.K,), 1.0 / self.K, device=dev,
                                        requires_grad=calibration_lr != 0)
        self.label_opt = torch.optim.Adam([self.mixture_label], lr=calibration_lr) \
            if calibration_lr != 0 else None
        # training.py:134-136: bootstrap permutations of the failure days, as day multiplicities
        g = torch.Generator().manual_seed(seed)
        D = failure_plan.n_days
        self.day_weights: List[torch.Tensor] = []
        for _ in range(self.K):
            draw = torch.randint(0, D, (D,), generator=g)
            self.day_weights.append(torch.bincount(draw, minlength=D).float().to(dev))
        self.status = torch.zeros((), dtype=torch.int32, device=dev)
        self._seed = seed * 1_000_003 + 17 * self.rank
        self.last = {}

    def _next_seed(self) -> int:
        self._seed += self.world * 101
        return self._seed

    def _elbo(self, label, plan, weights=None):
        return elbo_loss(self.guide(label), plan, self.scales, self.P, weights, self._next_seed(),
                         self.status, self.world)

    def step(self) -> torch.Tensor:
        dev = self.fplan.device
        # ---- mixture label (training.py:180-191) ----
        if self.calibrate and self.label_opt is not None:
            for _ in range(self.substeps):
                loss = self._elbo(self.mixture_label, self.fplan)
                # only the label is updated here: no gradient of the flow's parameters is formed
                (self.mixture_label.grad,) = torch.autograd.grad(loss, [self.mixture_label])
                if self.world > 1:
                    dist.all_reduce(self.mixture_label.grad, group=self.group)
                torch.nn.utils.clip_grad_norm_([self.mixture_label], self.grad_clip)
                self.label_opt.step()
        # ---- failure model (training.py:193-296) ----
        self.fp.zero_grad()
        if self.mixture_label.grad is not None:
            self.mixture_label.grad = None
        total = torch.zeros((), device=dev)
        if self.calibrate:
            fe = torch.zeros((), device=dev)
            for j in range(self.K):
                label = torch.zeros(self.K, device=dev)
                label[j] = 1.0
                fe = fe + self._elbo(label, self.fplan, self.day_weights[j])
            total = total + self.elbo_weight * fe / self.K
        else:
            total = total + self.elbo_weight * self._elbo(torch.ones(self.K, device=dev), self.fplan)
        if not self.exclude_nominal:
            total = total + self.elbo_weight * self._elbo(torch.zeros(self.K, device=dev), self.nplan)
        if self.calibrate:
            total = total + self.calibration_weight * self._elbo(self.mixture_label.detach(), self.fplan)
        with torch.cuda.nvtx.range("ba.calnf.backward") if dev.type == "cuda" else _null():
            total.backward()
        self.fp.all_reduce(self.group)
        norm = self.fp.clip_grad_norm_(self.grad_clip)
        self.opt.step()
        self.sched.step()
        self.last = {"loss": total.detach(), "grad_norm": norm}
        return total.detach()

    def check(self):
        """Raises if any particle-day of any step so far failed (synchronises)."""
        bits = int(self.status.item())
        if bits:
            SimulationPlan.check_status(torch.tensor([bits], dtype=torch.int32))
