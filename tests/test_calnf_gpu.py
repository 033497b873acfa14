"""GPU tests of the CalNF path: the fused spline kernels (ba_rqs_forward / ba_rqs_backward)
against the torch restatement, and a short CalNF training run on a top-4 synthetic network
(scripts/training.py:177-296 step structure) with the loss falling."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("bins", [4, 8, 12])
@pytest.mark.parametrize("M", [1, 255, 256, 5000])
def test_rqs_kernels_match_torch(bins, M):
    from bayesair_b200.flows import rqs, rqs_torch
    torch.manual_seed(bins * 1000 + M)
    dev = torch.device("cuda:0")
    x = ((torch.rand(M, device=dev) * 2 - 1) * 5.5).requires_grad_(True)
    p = (torch.randn(M, 3 * bins - 1, device=dev) * 1.5).requires_grad_(True)
    gy, gl = torch.randn(M, device=dev), torch.randn(M, device=dev)
    y, ld = rqs(x, p, bins, 5.0)
    (gx, gp) = torch.autograd.grad((y * gy + ld * gl).sum(), (x, p))
    x64, p64 = x.detach().double().requires_grad_(True), p.detach().double().requires_grad_(True)
    yr, lr = rqs_torch(x64, p64, bins, 5.0)
    (gxr, gpr) = torch.autograd.grad((yr * gy.double() + lr * gl.double()).sum(), (x64, p64))
    # fp32 against fp64: a point inside a very narrow bin (width down to 0.01) sees its knot
    # rounding amplified by 1 / width, so the bound on the WORST element is loose and the
    # bound on the typical element tight
    def check(a, b, typical, worst):
        err = (a.double() - b).abs() / (1.0 + b.abs())
        assert float(err.median()) <= typical and float(err.max()) <= worst, (float(err.median()), float(err.max()))
    check(y, yr, 2e-6, 1e-3)
    check(ld, lr, 5e-6, 5e-3)
    nx, npar = gxr.abs().max().clamp_min(1e-30), gpr.abs().max().clamp_min(1e-30)
    check(gx / nx, gxr / nx, 1e-5, 5e-3)
    check(gp / npar, gpr / npar, 1e-5, 5e-3)


def _network(n_airports, flights, days, seed):
    from bayesair_b200.compiler import day_spec_from_frame
    from bayesair_b200.synth import synth_day_frame
    frames = [synth_day_frame(n_airports, flights, d, seed=seed) for d in range(days)]
    specs = [day_spec_from_frame(frames[0])]
    specs += [day_spec_from_frame(f, specs[0].codes) for f in frames[1:]]
    return specs


def test_calnf_training_run_on_top4_network():
    from bayesair_b200.calnf import CalNFTrainer
    from bayesair_b200.flows import ConditionalSplineFlow
    from bayesair_b200.objective import n_latent
    from bayesair_b200.plan import SimulationPlan
    from bayesair_b200.synth import route_times
    torch.manual_seed(0)
    N, K = 4, 2
    dev = torch.device("cuda:0")
    fail = _network(N, 80, 3, seed=11)
    nom = _network(N, 80, 2, seed=12)
    fplan = SimulationPlan(fail, fail[0].codes, delta_t=0.2, device=dev)
    nplan = SimulationPlan(nom, fail[0].codes, delta_t=0.2, device=dev)
    n = n_latent(N)
    init = torch.zeros(n)
    init[:N] = np.log(0.5)
    init[N:2 * N] = np.log(0.05)               # off the data: service times too long
    init[2 * N:] = torch.log(torch.tensor(route_times(N, 11), dtype=torch.float32).clamp_min(0.5)).reshape(-1)
    guide = ConditionalSplineFlow(n, K, hidden_features=(32, 32), transforms=4, init_loc=init,
                                  init_scale=0.1).to(dev)
    trainer = CalNFTrainer(guide, fplan, nplan, torch.tensor([0.1, 0.1, 0.1], device=dev),
                           n_particles=64, n_permutations=K, lr=5e-3, calibration_lr=1e-2, seed=3)
    label0 = trainer.mixture_label.detach().clone()
    losses = [float(trainer.step()) for _ in range(30)]
    trainer.check()
    assert all(np.isfinite(losses))
    assert np.mean(losses[-5:]) < np.mean(losses[:5]), losses
    assert not torch.equal(label0, trainer.mixture_label.detach())       # the label is trained
    # the guide still is a proper density after training
    d = guide(trainer.mixture_label.detach())
    z, lq = d.rsample_and_log_prob((16,))
    assert torch.allclose(d.log_prob(z.detach()), lq.detach(), atol=2e-2, rtol=1e-3)
