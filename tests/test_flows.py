"""CPU tests of the conditional spline-flow guide (bayesair_b200/flows.py) and the flat
parameter buffer of the CalNF trainer: the surface scripts/training.py:184-294 and
scripts/utils.py:16-26 use, the spline's inverse / log-determinant, density consistency."""
import math

import torch

from bayesair_b200.flows import ConditionalSplineFlow, rqs_torch


def _randomise(flow, std=0.3):
    for layer in flow.layers:
        torch.nn.init.normal_(layer.out.weight, std=std)
        torch.nn.init.normal_(layer.out.bias, std=std)


def test_spline_is_monotone_invertible_and_logdet_is_its_derivative():
    torch.manual_seed(0)
    x = (torch.rand(400, 5, dtype=torch.float64) * 2 - 1) * 6.0       # includes the tails
    p = torch.randn(400, 5, 23, dtype=torch.float64) * 1.5
    xr = x.clone().requires_grad_(True)
    y, ld = rqs_torch(xr, p, 8, 5.0)
    (g,) = torch.autograd.grad(y.sum(), xr)
    assert (g > 0).all()
    assert torch.allclose(g.log(), ld, atol=1e-10)
    xi, ldi = rqs_torch(y.detach(), p, 8, 5.0, inverse=True)
    assert torch.allclose(xi, x, atol=1e-9)
    assert torch.allclose(ldi, -ld.detach(), atol=1e-9)
    outside = x.abs() >= 5.0
    assert torch.equal(y.detach()[outside], x[outside]) and (ld.detach()[outside] == 0).all()
    # zero parameters: the identity
    y0, ld0 = rqs_torch(x, torch.zeros_like(p), 8, 5.0)
    assert torch.allclose(y0, x, atol=1e-12) and torch.allclose(ld0, torch.zeros_like(ld0), atol=1e-12)


def test_flow_surface_and_density_consistency():
    torch.manual_seed(1)
    n, K = 24, 3
    flow = ConditionalSplineFlow(n, K, hidden_features=(16, 16), transforms=4,
                                 init_loc=torch.linspace(-4.0, 1.0, n), init_scale=0.3).double()
    _randomise(flow)
    label = torch.tensor([0.0, 1.0, 0.0], dtype=torch.float64)
    d = flow(label)
    z, lq = d.rsample_and_log_prob((64,))
    assert z.shape == (64, n) and lq.shape == (64,)
    assert torch.allclose(d.log_prob(z), lq, atol=1e-8)
    # a different label is a different distribution, the same label the same one
    assert not torch.allclose(flow(torch.tensor([1.0, 0, 0], dtype=torch.float64)).log_prob(z), lq)
    assert d.sample((3,)).requires_grad is False
    assert d.rsample((3,)).requires_grad
    z1, lq1 = d.rsample_and_log_prob()
    assert z1.shape == (n,) and lq1.shape == ()
    # reparameterised: gradients reach every conditioner
    lq.sum().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in flow.parameters())


def test_fresh_flow_is_the_fixed_affine_map_of_a_unit_gaussian():
    torch.manual_seed(2)
    n = 10
    loc, sc = torch.linspace(-3.0, 0.0, n), 0.5
    flow = ConditionalSplineFlow(n, 2, init_loc=loc, init_scale=sc)
    z, lq = flow(torch.zeros(2)).rsample_and_log_prob((2000,))
    u = (z - loc) / sc
    ref = (-0.5 * u * u).sum(-1) - 0.5 * n * math.log(2 * math.pi) - n * math.log(sc)
    assert torch.allclose(lq, ref, atol=1e-4)
    assert abs(float(u.mean())) < 0.05 and abs(float(u.std()) - 1.0) < 0.05


def test_density_integrates_to_one_in_two_dimensions():
    torch.manual_seed(3)
    flow = ConditionalSplineFlow(2, 1, hidden_features=(8,), transforms=2).double()
    _randomise(flow, 0.1)
    g = torch.linspace(-8, 8, 801, dtype=torch.float64)
    xx, yy = torch.meshgrid(g, g, indexing="ij")
    pts = torch.stack([xx.reshape(-1), yy.reshape(-1)], -1)
    with torch.no_grad():
        lp = flow(torch.ones(1, dtype=torch.float64)).log_prob(pts)
    mass = float(lp.exp().sum() * (g[1] - g[0]) ** 2)
    assert abs(mass - 1.0) < 2e-3


def test_flat_parameters_alias_the_module():
    from bayesair_b200.calnf import FlatParameters
    torch.manual_seed(4)
    flow = ConditionalSplineFlow(6, 2, hidden_features=(8, 8), transforms=2)
    _randomise(flow)
    before = [p.detach().clone() for p in flow.parameters()]
    fp = FlatParameters(flow)
    assert all(torch.equal(a, b) for a, b in zip(before, flow.parameters()))
    assert fp.flat.numel() == sum(p.numel() for p in flow.parameters())
    z, lq = flow(torch.tensor([1.0, 0.0])).rsample_and_log_prob((8,))
    (lq.sum() + (z ** 2).sum()).backward()
    # autograd accumulated straight into the flat buffer
    got = torch.cat([p.grad.reshape(-1) for p in flow.parameters()])
    assert torch.equal(got, fp.grad) and float(fp.grad.abs().sum()) > 0
    assert all(p.grad.data_ptr() >= fp.grad.data_ptr() for p in flow.parameters())
    opt = torch.optim.Adam([fp.parameter], lr=1e-2)
    opt.step()
    assert not torch.equal(before[0], next(flow.parameters()).detach())
    fp.zero_grad()
    assert float(torch.cat([p.grad.reshape(-1) for p in flow.parameters()]).abs().sum()) == 0.0
    n = fp.clip_grad_norm_(1.0)
    assert float(n) == 0.0
