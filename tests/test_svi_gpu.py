"""GPU tests of the fused mean-field SVI step (bayesair_b200/svi.py, csrc/ba_guide.cu):
the library's flat gradient buffer against framework autograd of the same objective
(scripts/wn/train.py:301-326), graph replay against eager stepping, and a short training
run."""
import math

import numpy as np
import pytest
import torch

from tests.util import Golden

pytestmark = pytest.mark.gpu


def _plan(g):
    from bayesair_b200.plan import SimulationPlan
    return SimulationPlan(g.specs, g.global_codes, delta_t=g.delta_t, device="cuda:0")


def _init_loc(g, N):
    L = g.latents
    tr = np.maximum(L["travel"].copy(), 0.25)
    np.fill_diagonal(tr, 1.0)
    return torch.tensor(np.concatenate([np.log(L["mean_turnaround"]), np.log(L["mean_service"]),
                                        np.log(tr).ravel()]).astype(np.float32))


@pytest.mark.parametrize("name", ["top10", "twoday5"])
def test_flat_gradient_matches_autograd(name):
    """[d loc | d log_scale | d scales | loss] written by ba_mf_backward equals what autograd
    gives for loss = -mean_p(prior + sum_d logp - log q) / flights with z = loc + sigma eps."""
    from bayesair_b200.function import simulate_log_joint
    from bayesair_b200.objective import map_to_latents, prior_log_prob
    from bayesair_b200.svi import FusedMeanFieldSVI
    g = Golden(name)
    plan = _plan(g)
    N, P = plan.n_airports, 48
    scales = torch.tensor([0.1, 0.12, 0.09], device=plan.device)
    svi = FusedMeanFieldSVI(plan, scales, P, init_loc=_init_loc(g, N), init_scale=0.07, seed=1234,
                            use_graph=False)
    n = svi.n
    z = torch.empty((P, n), device=plan.device)
    out = svi.local_gradient(z_out=z)
    torch.cuda.synchronize()
    plan.check_status(out["status"])
    flat = svi.flat.detach().cpu().double().numpy()

    loc = svi.loc.clone().requires_grad_(True)
    lsc = svi.log_scale.clone().requires_grad_(True)
    sc = scales.clone().requires_grad_(True)
    eps = ((z - loc) * torch.exp(-lsc)).detach()
    # the guide's draws are standard normal
    assert abs(float(eps.mean())) < 0.05 and abs(float(eps.std()) - 1.0) < 0.05
    zz = loc + torch.exp(lsc) * eps
    logq = (-0.5 * eps ** 2 - lsc - 0.5 * math.log(2.0 * math.pi)).sum(-1)
    lat = map_to_latents(zz, N)
    la, lr = plan.pack_latents(lat["mean_turnaround"], lat["mean_service"], lat["travel"])
    assert torch.allclose(la, svi.la, rtol=2e-6) and torch.allclose(lr, svi.lr_, rtol=2e-6)
    logp = simulate_log_joint(plan, la, lr, sc, None, seed=1234)
    assert torch.equal(logp.detach(), out["logp"])        # same Philox stream
    elbo = prior_log_prob(lat) + logp.sum(1) - logq
    loss = -elbo.mean() / plan.n_flights_total
    loss.backward()
    ref = np.concatenate([loc.grad.cpu().double().numpy(), lsc.grad.cpu().double().numpy(),
                          sc.grad.cpu().double().numpy(), [float(loss)]])
    scale = np.maximum(np.abs(ref), 1e-6 * np.abs(ref).max())
    err = np.abs(flat - ref) / scale
    assert err[:n].max() < 2e-3, ("loc", err[:n].max())
    assert err[n:2 * n].max() < 2e-3, ("log_scale", err[n:2 * n].max())
    assert err[2 * n:2 * n + 3].max() < 1e-3, ("scales", err[2 * n:2 * n + 3])
    assert err[2 * n + 3] < 1e-5, ("loss", flat[-1], ref[-1])


def test_graph_replay_equals_eager_and_trains():
    """The captured step replays bit-identically to eager stepping (device-side seed), the
    status flag stays clear, and the loss falls over a short run."""
    from bayesair_b200.svi import FusedMeanFieldSVI
    g = Golden("top10")
    plan = _plan(g)
    N = plan.n_airports
    scales = torch.tensor([0.1, 0.1, 0.1], device=plan.device)
    init = _init_loc(g, N) + 0.3          # start off the generating latents
    runs = []
    for use_graph in (False, True):
        svi = FusedMeanFieldSVI(plan, scales, 64, init_loc=init, init_scale=0.05, lr=2e-2, seed=7,
                                use_graph=use_graph)
        losses = [svi.step().clone() for _ in range(40)]
        torch.cuda.synchronize()
        svi.check()
        assert use_graph == (svi._graph is not None)
        runs.append((torch.stack(losses).cpu(), svi.theta.detach().cpu().clone()))
    (l0, t0), (l1, t1) = runs
    assert torch.equal(l0, l1)
    assert torch.equal(t0, t1)
    assert float(l0[-5:].mean()) < float(l0[:5].mean())


# ---------------------------------------------------------------- call site (ii) ----------
def _golden_states(g):
    from tests.test_cuda_parity import _states_from_golden
    return _states_from_golden(g)


def test_scaled_model_trace_matches_reference_total():
    """poutine.scale(model, 1/days) -- scripts/wn/wn_network.py:285-287 -- scales every site,
    the fused factor included: the trace's log_prob_sum is the reference-generated total of
    the golden divided by the number of days."""
    import bayesair_b200 as ba
    from bayesair_b200.ppl import poutine, pyro
    g = Golden("twoday5")
    states = _golden_states(g)
    dev = torch.device("cuda:0")
    codes = g.global_codes
    data = {}
    for i, c in enumerate(codes):
        data[f"{c}_mean_turnaround_time"] = torch.tensor(g.latents["mean_turnaround"][i], device=dev)
        data[f"{c}_mean_service_time"] = torch.tensor(g.latents["mean_service"][i], device=dev)
        for j, c2 in enumerate(codes):
            if i != j:
                data[f"travel_time_{c}_{c2}"] = torch.tensor(g.latents["travel"][i, j], device=dev)
    pyro.clear_param_store()
    for nme, v in zip(("runway_use_time_std_dev", "travel_time_variation",
                       "turnaround_time_variation"), g.scales):
        pyro.param(nme, torch.tensor(v, device=dev), constraint=torch.distributions.constraints.positive)
    days = len(states)
    model = poutine.scale(ba.air_traffic_network_model, scale=1.0 / days)
    with ba.flight_noise(torch.tensor(g.noise[None], device=dev)):
        tr = poutine.trace(poutine.condition(model, data=data)).get_trace(states, g.delta_t, device=dev)
    assert abs(float(tr.log_prob_sum()) - g.ref_total / days) <= 1e-4 * abs(g.ref_total / days)


@pytest.mark.parametrize("guide_name", ["AutoNormal", "AutoMultivariateNormal"])
def test_svi_with_autoguide_over_the_scaled_model(guide_name):
    """The whole of wn_network.py:285-302 against the drop-in model: scale handler, autoguide
    over the global latent sites, ClippedAdam, Trace_ELBO, ``svi.step(states, dt)`` with a
    positional delta_t.  The loss of a step equals the ELBO assembled by hand from the same
    draws, and it falls over a short run."""
    import bayesair_b200 as ba
    from bayesair_b200.ppl import USING_REAL_PYRO, poutine, pyro
    g = Golden("twoday5")
    states = _golden_states(g)
    dev = torch.device("cuda:0")
    days = len(states)
    pyro.clear_param_store()
    torch.manual_seed(0)
    model = poutine.scale(ba.air_traffic_network_model, scale=1.0 / days)

    def model_on_device(states, dt):
        return model(states, dt, device=dev)

    guide = getattr(pyro.infer.autoguide, guide_name)(model_on_device)
    elbo = pyro.infer.Trace_ELBO(num_particles=8, vectorize_particles=not USING_REAL_PYRO)
    svi = pyro.infer.SVI(model_on_device, guide, pyro.optim.ClippedAdam({"lr": 0.02, "lrd": 0.999}), elbo)
    with ba.flight_noise(seed=5):
        first = svi.step(states, g.delta_t)
    assert np.isfinite(first)
    # one step's loss, by hand: guide trace -> replayed model trace -> scaled sums
    if not USING_REAL_PYRO:
        torch.manual_seed(123)
        with ba.flight_noise(seed=9):
            loss = float(elbo.differentiable_loss(model_on_device, guide, states, g.delta_t))
        torch.manual_seed(123)
        pyro._PARTICLE_SHAPE = (8,)
        try:
            gtr = poutine.trace(guide).get_trace(states, g.delta_t)
            with ba.flight_noise(seed=9):
                mtr = poutine.trace(poutine.replay(model_on_device, trace=gtr)).get_trace(states, g.delta_t)
        finally:
            pyro._PARTICLE_SHAPE = ()
        lp_model = sum((s["fn"].log_prob(s["value"]) / days).sum() for s in mtr.nodes.values()
                       if s["type"] == "sample")
        lp_guide = sum(s["fn"].log_prob(s["value"]).sum() for s in gtr.nodes.values()
                       if s["type"] == "sample")
        assert abs(loss - float(-(lp_model - lp_guide) / 8)) <= 1e-4 * abs(loss)
    losses = []
    for i in range(60):
        with ba.flight_noise(seed=100 + i):
            losses.append(svi.step(states, g.delta_t))
    assert np.mean(losses[-10:]) < np.mean(losses[:10])
    med = guide.median(states, g.delta_t)
    assert set(med) >= {f"{c}_mean_service_time" for c in g.global_codes}


def test_day_sharded_ranks_sum_to_the_unsharded_gradient():
    """Days sharded over ranks (parallel.shard_particle_days): two virtual ranks that own one
    day each of the same particles -- same seed, same particle offset, prior_weight 1/2 --
    produce flat buffers whose sum (what the all-reduce computes) is the unsharded buffer."""
    from bayesair_b200.plan import SimulationPlan
    from bayesair_b200.svi import FusedMeanFieldSVI
    g = Golden("twoday5")
    full = _plan(g)
    N, P = full.n_airports, 32
    scales = torch.tensor([0.1, 0.1, 0.1], device=full.device)
    init = _init_loc(g, N)
    ref = FusedMeanFieldSVI(full, scales, P, init_loc=init, init_scale=0.06, seed=77, use_graph=False)
    ref.local_gradient()
    torch.cuda.synchronize()
    total = torch.zeros_like(ref.flat)
    for d in range(2):
        plan_d = SimulationPlan([g.specs[d]], g.global_codes, delta_t=g.delta_t, device="cuda:0")
        part = FusedMeanFieldSVI(plan_d, scales, P, init_loc=init, init_scale=0.06, seed=77,
                                 use_graph=False, particle_offset=0, n_particles_total=P,
                                 n_flights_total=full.n_flights_total, prior_weight=0.5,
                                 day_offset=d)
        part.local_gradient()
        torch.cuda.synchronize()
        total += part.flat
    a, b = total.cpu().double().numpy(), ref.flat.cpu().double().numpy()
    scale = np.maximum(np.abs(b), 1e-6 * np.abs(b).max())
    assert (np.abs(a - b) / scale).max() < 1e-4, (np.abs(a - b) / scale).max()


def test_model_routes_value_gradients_to_a_per_flight_guide():
    """A guide over the per-flight sites hands their values to the drop-in model through
    flight_noise(values, noise_is_value=True); the factor's gradient w.r.t. those values
    equals the reference's autograd (value-mode golden)."""
    import bayesair_b200 as ba
    from bayesair_b200.ppl import poutine, pyro
    from tests.test_host_side import _value_golden, value_grad_error
    g, v = Golden("twoday5"), _value_golden("twoday5")
    states = _golden_states(g)
    dev = torch.device("cuda:0")
    codes = g.global_codes
    data = {}
    for i, c in enumerate(codes):
        data[f"{c}_mean_turnaround_time"] = torch.tensor(g.latents["mean_turnaround"][i], device=dev)
        data[f"{c}_mean_service_time"] = torch.tensor(g.latents["mean_service"][i], device=dev)
        for j, c2 in enumerate(codes):
            if i != j:
                data[f"travel_time_{c}_{c2}"] = torch.tensor(g.latents["travel"][i, j], device=dev)
    pyro.clear_param_store()
    for nme, s in zip(("runway_use_time_std_dev", "travel_time_variation",
                       "turnaround_time_variation"), g.scales):
        pyro.param(nme, torch.tensor(s, device=dev), constraint=torch.distributions.constraints.positive)
    vals = torch.tensor(np.nan_to_num(v["values"], nan=1.0)[None], device=dev, requires_grad=True)
    with ba.flight_noise(vals, noise_is_value=True):
        tr = poutine.trace(poutine.condition(ba.air_traffic_network_model, data=data)).get_trace(
            states, g.delta_t, device=dev)
    factor = tr.nodes[ba.FACTOR_NAME]["fn"].log_prob(None).sum()
    assert abs(float(factor) - float(v["total"])) <= 1e-4 * abs(float(v["total"]))
    factor.backward()
    assert value_grad_error(vals.grad[0].cpu().numpy(), v["grad_values"], v["values"]) <= 1e-3


def test_exposed_flight_sites_reproduce_the_reference_trace_and_take_a_guide():
    """expose_flight_sites=True: the four latent per-flight sites of the reference are real
    sites again.  Conditioned on the golden's values (globals and per-flight), the trace's
    log_prob_sum is the reference's total and its gradient w.r.t. the latents the reference's
    value-mode autograd; an autoguide then covers EVERY latent site, as
    AutoMultivariateNormal(model) does for the reference (wn_network.py:285-302)."""
    import bayesair_b200 as ba
    from bayesair_b200.ppl import USING_REAL_PYRO, poutine, pyro
    from tests.test_host_side import _value_golden
    from tests.util import rel_err
    g, v = Golden("tiny4"), _value_golden("tiny4")
    states = _golden_states(g)
    dev = torch.device("cuda:0")
    codes = g.global_codes
    ms = torch.tensor(g.latents["mean_service"], device=dev, requires_grad=True)
    data = {}
    for i, c in enumerate(codes):
        data[f"{c}_mean_turnaround_time"] = torch.tensor(g.latents["mean_turnaround"][i], device=dev)
        data[f"{c}_mean_service_time"] = ms[i]
        for j, c2 in enumerate(codes):
            if i != j:
                data[f"travel_time_{c}_{c2}"] = torch.tensor(g.latents["travel"][i, j], device=dev)
    kinds = ("departure_service_time", "travel_time", "arrival_service_time", "turnaround_time")
    n_flight_sites = 0
    for d, spec in enumerate(g.specs):
        for f in range(spec.n_flights):
            if spec.cancel_obs[f] == 1:
                continue
            for col, kind in enumerate(kinds):
                data[f"day{d}_{spec.flight_names[f]}_{kind}"] = torch.tensor(float(v["values"][d, f, col]), device=dev)
                n_flight_sites += 1
    pyro.clear_param_store()
    for nme, s in zip(("runway_use_time_std_dev", "travel_time_variation",
                       "turnaround_time_variation"), g.scales):
        pyro.param(nme, torch.tensor(s, device=dev), constraint=torch.distributions.constraints.positive)

    def model(states, dt):
        return ba.air_traffic_network_model(states, dt, device=dev, expose_flight_sites=True)

    tr = poutine.trace(poutine.condition(model, data=data)).get_trace(states, g.delta_t)
    total = tr.log_prob_sum()
    assert abs(float(total) - g.ref_total) <= 1e-4 * abs(g.ref_total)
    flight_nodes = [n for n in tr.nodes if n.startswith("day")]
    assert len(flight_nodes) == n_flight_sites
    per_flight = sum(tr.nodes[n]["fn"].log_prob(tr.nodes[n]["value"]).sum() for n in flight_nodes) \
        + tr.nodes[ba.FACTOR_NAME]["fn"].log_prob(None).sum()
    (gm,) = torch.autograd.grad(per_flight, [ms])
    assert rel_err(gm.cpu().numpy(), v["grad_mean_service"], floor=1e-4) <= 1e-3
    if USING_REAL_PYRO:
        return
    # a guide over every latent site: globals and the 4 per live flight
    pyro.clear_param_store()
    torch.manual_seed(0)
    guide = pyro.infer.autoguide.AutoNormal(model, init_scale=0.02)
    svi = pyro.infer.SVI(model, guide, pyro.optim.ClippedAdam({"lr": 0.01}),
                         pyro.infer.Trace_ELBO(num_particles=4, vectorize_particles=True))
    losses = [svi.step(states, g.delta_t) for _ in range(12)]
    assert np.isfinite(losses).all()
    assert guide.latent_dim == n_flight_sites + 2 * len(codes) + len(codes) * (len(codes) - 1)
    assert np.mean(losses[-4:]) < np.mean(losses[:4])
