"""CPU tests of the host side: schedule compiler orderings, the C-ABI library
(loads, exports every symbol of include/bayesair_b200.h, refuses to run without a
GPU), the PPL surface, and the emulated kernel against golden vectors + oracle."""
import os
import re

import numpy as np
import pytest
import torch

from tests.util import (Golden, PredictiveGolden, golden_names, predictive_names, rel_err,
                        std_noise)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---------------------------------------------------------------- C ABI -----
def test_library_exports_every_declared_symbol():
    from bayesair_b200 import _capi
    lib = _capi.load_library()
    header = open(os.path.join(ROOT, "include", "bayesair_b200.h")).read()
    declared = set(re.findall(r"\b(ba_[a-z_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_capi.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.ba_version() == 200


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_gpu():
    import bayesair_b200 as ba
    from bayesair_b200._capi import BayesAirLibraryError
    from bayesair_b200.synth import synth_day_frame
    fl, ap = ba.parse_schedule(synth_day_frame(4, 60))
    st = ba.NetworkState({a.code: a for a in ap}, fl)
    with pytest.raises(BayesAirLibraryError):
        ba.air_traffic_network_model([st], 0.2)
    with pytest.raises(BayesAirLibraryError):
        ba.air_traffic_network_model([st], 0.2, device=torch.device("cpu"))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "bayesair_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f
                assert "ba_oracle" not in src, f
                assert "hostemu" not in src or f in ("ba_sim.cuh", "ba_scan2.cuh"), f


# ------------------------------------------------------- schedule compiler --
def test_compiler_orderings():
    import bayesair_b200 as ba
    from bayesair_b200.compiler import compile_states, day_spec_from_frame, used_routes
    from bayesair_b200.synth import synth_day_frame
    frames = [synth_day_frame(6, 120, d, seed=3, cancel_frac=0.1) for d in range(2)]
    states = []
    for df in frames:
        # shuffle rows: NetworkState must restore stable scheduled-departure order
        fl, ap = ba.parse_schedule(df)
        states.append(ba.NetworkState({a.code: a for a in ap}, fl))
    codes, specs = compile_states(states)
    assert codes == list(states[0].airports.keys())
    for df, s in zip(frames, specs):
        assert np.all(np.diff(s.sched_dep) >= 0)
        v = day_spec_from_frame(df, codes)
        for k in ("sched_dep", "act_dep", "act_arr", "cancel_obs", "origin", "dest", "airport_global"):
            assert np.array_equal(getattr(s, k), getattr(v, k), equal_nan=True), k
        assert s.flight_names == v.flight_names
        assert np.isnan(s.act_dep[s.cancel_obs == 1]).all()
    # day 1 lists airports in its own first-appearance order, mapped to day-0 indices
    assert sorted(specs[1].airport_global.tolist()) == list(range(6))
    assert [codes[g] for g in specs[1].airport_global] == specs[1].codes
    r = used_routes(specs, 6)
    assert len(r) == len(set(r.tolist())) and (r // 6 != r % 6).all()


def test_stable_sort_keeps_tie_order():
    import bayesair_b200 as ba
    t = lambda x: torch.tensor(x)  # noqa: E731
    fl = [ba.Flight(str(i), "A", "B", t(s), t(s + 1.0), t(0.0), actual_departure_time=t(s),
                    actual_arrival_time=t(s + 1)) for i, s in enumerate([7.0, 6.0, 7.0, 6.0, 5.0])]
    assert str(fl[0]) == "0_A_B"
    st = ba.NetworkState({"A": ba.Airport("A"), "B": ba.Airport("B")}, fl)
    assert [f.flight_number for f in st.pending_flights] == ["4", "1", "3", "0", "2"]
    assert not st.complete


# -------------------------------------------------------------- PPL surface --
def test_mini_ppl_handlers():
    from bayesair_b200.ppl import USING_REAL_PYRO, dist, poutine, pyro
    pyro.clear_param_store()

    def model():
        s = pyro.param("s", torch.tensor(0.1), constraint=dist.constraints.positive)
        a = pyro.sample("a", dist.Gamma(torch.tensor(1.5), torch.tensor(10.0)))
        pyro.factor("f", -(a / s) ** 2)
        return a

    a0 = torch.tensor(0.3, requires_grad=True)
    tr = poutine.trace(poutine.condition(model, data={"a": a0})).get_trace()
    expect = torch.distributions.Gamma(1.5, 10.0).log_prob(a0) - (a0 / 0.1) ** 2
    assert abs(float(tr.log_prob_sum()) - float(expect)) < 1e-4
    tr2 = poutine.trace(poutine.scale(poutine.condition(model, data={"a": a0}), scale=0.5)).get_trace()
    assert abs(float(tr2.log_prob_sum()) - 0.5 * float(expect)) < 1e-4
    assert abs(float(pyro.param("s")) - 0.1) < 1e-7
    if not USING_REAL_PYRO:
        with pytest.raises(RuntimeError):
            poutine.trace(lambda: (model(), model())).get_trace()


# ---------------------------------------------- emulated kernel (CPU, g++) --
@pytest.mark.parametrize("exact", [False, True])
@pytest.mark.parametrize("name", [n for n in golden_names() if n != "wn100"])
def test_emulated_kernel_vs_golden(name, exact):
    from tests.hostemu import EmuPlan
    g = Golden(name)
    plan = EmuPlan(g.specs, len(g.global_codes), g.delta_t, g.max_t, g.include_cancellations)
    lat = {k: v[None] for k, v in g.latents.items()}
    o = plan.forward(lat, g.scales, g.noise[None], exact=exact, shuffle_seed=3)
    assert (o["status"] == 0).all()
    ref = sum(g.ref_class.values())
    assert abs(float(o["logp"].astype(np.float64).sum()) - ref) <= 5e-6 * abs(ref)
    for k, v in g.ref_grads.items():
        assert rel_err(o["grads"][k][0], v) <= 1e-4, k
    assert rel_err(o["grads"]["scales"][0], g.ref_grad_scales) <= 1e-4


@pytest.mark.parametrize("canc", [False, True])
def test_emulated_kernel_decisions_vs_oracle(canc):
    """Pop tick of every flight (the queue decisions) identical to the oracle for
    several particles and lane interleavings, incl. a starved-aircraft regime."""
    from bayesair_b200.compiler import day_spec_from_frame
    from bayesair_b200.synth import synth_day_frame, synth_latents
    from oracle import ba_oracle
    from tests.hostemu import EmuPlan
    N, F, P = 8, 320, 6
    frames = [synth_day_frame(N, F, d, seed=77, cancel_frac=0.05 if canc else 0.0) for d in range(2)]
    specs = [day_spec_from_frame(frames[0])]
    specs.append(day_spec_from_frame(frames[1], specs[0].codes))
    lat = synth_latents(N, P, seed=9, service_scale=0.06, spread=0.6,
                        include_cancellations=canc, n_aircraft_mean=2.5)
    noise = std_noise(P, 2, F, seed=5)
    sc = (0.1, 0.1, 0.1)
    plan = EmuPlan(specs, N, 0.2, 30.0, canc)
    for shuffle in (1, 2):
        o = plan.forward(lat, sc, noise, shuffle_seed=shuffle, want_pop=True,
                         replay_waits=(shuffle == 2))
        assert (o["status"] == 0).all()
        for p in range(P):
            for d in range(2):
                r = ba_oracle.run_day(specs[d], {k: v[p] for k, v in lat.items()}, sc,
                                      noise[p, d], delta_t=0.2, include_cancellations=canc)
                assert r["n_ticks"] == o["ticks"][p, d]
                assert np.array_equal(o["pop_tick"][p, d], r["pop_tick"])
                scale = sum(abs(v) for v in r["class_sums"].values())
                assert abs(r["logp"] - o["logp"][p, d]) <= 2e-6 * scale


def test_emulated_ring_spill(monkeypatch):
    """Starved shared-memory rings (capacity 1) spill to the global ring and refill:
    same decisions and log-prob as the exact-capacity variant, no overflow."""
    from tests.hostemu import EmuPlan
    g = Golden("congested6")
    monkeypatch.setenv("BA_Q_MIN", "1"); monkeypatch.setenv("BA_Q_DIV", "1e9")
    plan = EmuPlan(g.specs, len(g.global_codes), g.delta_t, g.max_t, False)
    lat = {k: v[None] for k, v in g.latents.items()}
    spill = plan.forward(lat, g.scales, g.noise[None], rerun_overflow=False, want_pop=True)
    exact = plan.forward(lat, g.scales, g.noise[None], exact=True, want_pop=True)
    assert (spill["status"] == 0).all()
    assert np.array_equal(spill["pop_tick"], exact["pop_tick"])
    assert np.array_equal(spill["logp"], exact["logp"])
    replay = plan.forward(lat, g.scales, g.noise[None], rerun_overflow=False, want_pop=True,
                          replay_waits=True)
    assert (replay["status"] == 0).all() and np.array_equal(replay["pop_tick"], exact["pop_tick"])


def _burst_day(n_origins=20, per_origin=20):
    """20 airports x 20 flights to one hub, all departing within seconds of each other: with a
    tight travel-time scale the ~400 arrivals fall into one or two ticks and overflow the
    shared due buffer."""
    from bayesair_b200.compiler import DaySpec
    f32 = np.float32
    n = n_origins * per_origin
    origin = np.repeat(np.arange(1, n_origins + 1), per_origin).astype(np.int32)
    order = np.argsort(np.tile(np.arange(per_origin), n_origins), kind="stable")
    origin = origin[order]                      # interleave the origins
    sched = np.full(n, 6.0, f32)
    dep = (6.0 + 0.0001 * np.arange(n)).astype(f32)
    codes = ["HUB"] + [f"S{i:02d}" for i in range(n_origins)]
    return DaySpec(codes, np.arange(n_origins + 1, dtype=np.int32), sched, dep,
                   (dep + 2.0).astype(f32), np.zeros(n, np.int8), origin,
                   np.zeros(n, np.int32), [f"{i}_{codes[o]}_HUB" for i, o in enumerate(origin)])


def test_emulated_overflow_rerun():
    """More flights due in one tick than the shared due buffer holds: the fast variant
    flags BA_PD_OVERFLOW and the re-run with exact-capacity lists is transparent."""
    from oracle import ba_oracle
    from tests.hostemu import EmuPlan
    spec = _burst_day()
    N = spec.n_airports
    plan = EmuPlan([spec], N, 0.2, 30.0, False)
    lat = {"mean_turnaround": np.full((1, N), 0.5, np.float32),
           "mean_service": np.full((1, N), 0.01, np.float32),
           "travel": np.full((1, N, N), 2.0, np.float32)}
    noise = std_noise(1, 1, spec.n_flights, seed=8)
    sc = (0.1, 0.004, 0.1)
    flagged = plan.forward(lat, sc, noise, rerun_overflow=False)
    assert (flagged["status"] & 2).all()
    fixed = plan.forward(lat, sc, noise)
    ref = ba_oracle.run_batch([spec], lat, sc, noise, delta_t=0.2)
    assert (fixed["status"] == 0).all()
    np.testing.assert_allclose(fixed["logp"], ref["logp"], rtol=5e-6)


# ------------------------------ emulated ring-less scan (ba_scan2.cuh, g++) --
@pytest.mark.parametrize("name", [n for n in golden_names()
                                  if n not in ("cancel4", "cancel5", "cancel2day")])
def test_emulated_scan2_vs_golden(name):
    """The one-warp ring-less scan (the kernel the training configurations run) against the
    reference-generated fixtures: log-prob, every gradient element, and the same queue
    decisions as the exact-list kernel."""
    from tests.hostemu import EmuPlan
    g = Golden(name)
    plan = EmuPlan(g.specs, len(g.global_codes), g.delta_t, g.max_t, False)
    assert plan.scan2_ok
    lat = {k: v[None] for k, v in g.latents.items()}
    ex = plan.forward(lat, g.scales, g.noise[None], exact=True, want_pop=True)
    for shuffle in (0, 3):
        o = plan.forward(lat, g.scales, g.noise[None], scan2=True, shuffle_seed=shuffle, want_pop=True)
        assert (o["status"] == 0).all()
        ref = sum(g.ref_class.values())
        assert abs(float(o["logp"].astype(np.float64).sum()) - ref) <= 5e-6 * abs(ref)
        for k, v in g.ref_grads.items():
            assert rel_err(o["grads"][k][0], v) <= 1e-4, k
        assert rel_err(o["grads"]["scales"][0], g.ref_grad_scales) <= 1e-4
        assert np.array_equal(o["pop_tick"], ex["pop_tick"])
        assert np.array_equal(o["ticks"], ex["ticks"])
    # without diagnostics (what the benchmark runs): same bars
    o = plan.forward(lat, g.scales, g.noise[None], scan2=True, want_pop=False)
    assert (o["status"] == 0).all()
    assert abs(float(o["logp"].astype(np.float64).sum()) - ref) <= 5e-6 * abs(ref)
    for k, v in g.ref_grads.items():
        assert rel_err(o["grads"][k][0], v) <= 1e-4, ("no diagnostics", k)
    assert rel_err(o["grads"]["scales"][0], g.ref_grad_scales) <= 1e-4


SCAN2_CASES = [  # N, [F/day], P, mean service, spread
    (8, [320, 300], 4, 0.06, 0.6),       # one round, two days of different length
    (40, [900], 3, 0.15, 0.5),           # two rounds, saturated hubs: late boxes, replays
    (70, [1500, 1400], 2, 0.05, 0.5),    # three rounds
    (100, [3876], 2, 0.012, 0.3),        # the benchmark's day, nominal
    (100, [3876], 1, 0.12, 0.3),         # ... and with every hub saturated (671 ticks)
    (130, [2600], 2, 0.03, 0.5),         # five rounds
]


@pytest.mark.parametrize("case", SCAN2_CASES, ids=lambda c: f"N{c[0]}F{c[1][0]}s{c[3]}")
def test_emulated_scan2_vs_oracle(case):
    """Queue decisions (pop tick of every flight), tick counts, log-probs and gradients of the
    ring-less scan equal to the C oracle, for shuffled lane orders inside the rounds."""
    from bayesair_b200.compiler import day_spec_from_frame
    from bayesair_b200.synth import synth_day_frame, synth_latents
    from oracle import ba_oracle
    from tests.hostemu import EmuPlan
    N, Fs, P, svc, spread = case
    frames = [synth_day_frame(N, F, d, seed=50 + N) for d, F in enumerate(Fs)]
    specs = [day_spec_from_frame(frames[0])]
    specs += [day_spec_from_frame(f, specs[0].codes) for f in frames[1:]]
    lat = synth_latents(N, P, seed=N + 1, service_scale=svc, spread=spread)
    noise = std_noise(P, len(Fs), max(Fs), seed=N)
    sc = (0.09, 0.11, 0.1)
    ref = ba_oracle.run_batch(specs, lat, sc, noise, delta_t=0.2, n_threads=4)
    plan = EmuPlan(specs, N, 0.2, 30.0, False)
    for shuffle in (1, 2):
        o = plan.forward(lat, sc, noise, scan2=True, shuffle_seed=shuffle, want_pop=True)
        assert (o["status"] == 0).all()
        assert np.array_equal(o["ticks"], ref["n_ticks"])
        np.testing.assert_allclose(o["logp"], ref["logp"], rtol=5e-6, atol=5e-6 * np.abs(ref["logp"]).max())
        for k, v in ref["grads"].items():
            if k in o["grads"]:
                assert rel_err(o["grads"][k], v) <= 1e-4, k
        for d, spec in enumerate(specs):
            r = ba_oracle.run_day(spec, {k: v[0] for k, v in lat.items()}, sc, noise[0, d],
                                  delta_t=0.2, want_grad=False)
            assert np.array_equal(o["pop_tick"][0, d, :spec.n_flights], r["pop_tick"])


def test_emulated_scan2_value_mode_and_philox():
    """noise_is_value (site values fixed by a guide) and in-kernel Philox noise on the
    ring-less scan agree with the exact-list kernel."""
    from tests.hostemu import EmuPlan
    g = Golden("twoday5")
    plan = EmuPlan(g.specs, len(g.global_codes), g.delta_t, g.max_t, False)
    lat = {k: v[None] for k, v in g.latents.items()}
    vals = np.nan_to_num(g.ref_values, nan=1.0)[None]
    a = plan.forward(lat, g.scales, vals, value_mode=True, scan2=True)
    b = plan.forward(lat, g.scales, vals, value_mode=True, exact=True)
    ref_sum = sum(g.ref_class.values())
    assert abs(float(a["logp"].sum()) - ref_sum) <= 5e-6 * abs(ref_sum)
    for k in b["grads"]:
        assert rel_err(a["grads"][k], b["grads"][k]) <= 1e-5, k
    p1 = plan.forward(lat, g.scales, None, seed=11, scan2=True)
    p2 = plan.forward(lat, g.scales, plan.philox_noise(1, 11), scan2=True)
    assert np.array_equal(p1["logp"], p2["logp"])


def test_emulated_scan2_bucket_overflow_reruns_exact():
    """An arrival beyond the shared tick buckets flags BA_PD_OVERFLOW and the particle-day is
    re-run with the exact lists (same answer as the oracle)."""
    from oracle import ba_oracle
    from tests.hostemu import EmuPlan
    g = Golden("tiny4")
    plan = EmuPlan(g.specs, len(g.global_codes), 0.02, g.max_t, False)   # 50 ticks per hour
    lat = {k: v[None] for k, v in g.latents.items()}
    flagged = plan.forward(lat, g.scales, g.noise[None], scan2=True, rerun_overflow=False)
    assert (flagged["status"] & 2).all()
    fixed = plan.forward(lat, g.scales, g.noise[None], scan2=True)
    ref = ba_oracle.run_batch(g.specs, lat, g.scales, g.noise[None], delta_t=0.02)
    assert (fixed["status"] == 0).all()
    np.testing.assert_allclose(fixed["logp"], ref["logp"], rtol=5e-6)


def test_clock_matches_torch_inplace_add():
    """The kernel's clock (sequential fp32 adds, model.py:158-161) equals torch's
    in-place `t += delta_t` on an fp32 tensor, tick for tick."""
    t = torch.tensor(0.0)
    c = np.float32(0.0)
    for k in range(400):
        t += 0.2
        c = np.float32(c + np.float32(0.2))
        assert float(t) == float(c)
    assert abs(float(c) - 80.0) > 1e-5     # i.e. NOT k * dt


# ------------------------------------------------------ batched objective ---
def test_latent_layout_and_priors_match_reference_sites():
    """map_to_latents follows scripts/wn/train.py:244-297 and prior_log_prob equals the sum
    of the model's global-site priors (bayes_air/model.py:58-113)."""
    from bayesair_b200.objective import map_to_latents, n_latent, prior_log_prob
    N, P = 5, 3
    for canc in (False, True):
        torch.manual_seed(1)
        z = 0.3 * torch.randn(P, n_latent(N, canc)) - 1.0
        lat = map_to_latents(z, N, canc)
        assert torch.equal(lat["mean_turnaround"], torch.exp(z[:, :N]))
        assert torch.equal(lat["mean_service"], torch.exp(z[:, N:2 * N]))
        off = 4 * N if canc else 2 * N
        assert torch.equal(lat["travel"][:, 1, 2], torch.exp(z[:, off + 1 * N + 2]))
        td = torch.distributions
        ref = td.Gamma(1.0, 2.0).log_prob(lat["mean_turnaround"]).sum(1)
        ref = ref + td.Gamma(1.5, 10.0).log_prob(lat["mean_service"]).sum(1)
        mask = ~torch.eye(N, dtype=torch.bool)
        ref = ref + (td.Gamma(4.0, 1.25).log_prob(lat["travel"]) * mask).sum((1, 2))
        if canc:
            assert torch.equal(lat["log_initial_aircraft"], z[:, 2 * N:3 * N])
            ref = ref + td.Normal(0.0, 1.0).log_prob(lat["log_initial_aircraft"]).sum(1)
            ref = ref + td.Normal(-3.0, 1.0).log_prob(lat["base_cancel_logprob"]).sum(1)
        assert torch.allclose(prior_log_prob(lat, canc), ref, rtol=1e-5, atol=1e-4)


def test_mean_field_guide_surface():
    from bayesair_b200.objective import MeanFieldGuide
    g = MeanFieldGuide(7)
    z, lp = g.rsample_and_log_prob((5,))
    assert z.shape == (5, 7) and lp.shape == (5,)
    assert torch.allclose(g.log_prob(z), lp, atol=1e-5)
    assert g.sample((2,)).shape == (2, 7) and not g.sample((2,)).requires_grad


# ------------------------------------------------ posterior-predictive mode --
@pytest.mark.parametrize("name", predictive_names())
def test_emulated_predictive_vs_golden(name):
    """obs=None everywhere: the emulated kernel reproduces the reference's simulated
    cancellations exactly and its simulated times within 2 ulp, for two interleavings."""
    from tests.hostemu import EmuPlan
    g = PredictiveGolden(name)
    plan = EmuPlan(g.specs, len(g.global_codes), g.delta_t, g.max_t, g.include_cancellations)
    lat = {k: v[None] for k, v in g.latents.items()}
    for shuffle in (1, 4):
        o = plan.predict(lat, g.scales, g.noise[None], g.obs_noise[None], shuffle_seed=shuffle)
        assert (o["status"] == 0).all()
        for d in range(g.n_days):
            g.check(o["sim"][0, d], d)


def test_emulated_predictive_vs_oracle_batch():
    from bayesair_b200.compiler import day_spec_from_frame
    from bayesair_b200.synth import synth_day_frame, synth_latents
    from oracle import ba_oracle
    from tests.hostemu import EmuPlan
    N, F, P = 7, 260, 5
    spec = day_spec_from_frame(synth_day_frame(N, F, 0, seed=55))
    spec.act_dep[::2] = np.nan          # half the flights keep their observations
    spec.act_arr[::3] = np.nan
    spec.cancel_obs[:] = -1
    lat = synth_latents(N, P, seed=6, service_scale=0.05, spread=0.5, include_cancellations=True,
                        n_aircraft_mean=3.0)
    noise = std_noise(P, 1, F, seed=1)
    rng = np.random.default_rng(2)
    n2 = np.zeros((P, 1, F, 4), np.float32)
    n2[..., 0] = rng.random((P, 1, F)); n2[..., 1:3] = rng.standard_normal((P, 1, F, 2))
    sc = (0.1, 0.1, 0.1)
    ref = ba_oracle.run_batch([spec], lat, sc, noise, delta_t=0.2, include_cancellations=True,
                              want_grad=False, obs_noise=n2, want_sim=True)
    plan = EmuPlan([spec], N, 0.2, 30.0, True)
    o = plan.predict(lat, sc, noise, n2, shuffle_seed=3)
    assert (o["status"] == 0).all() and (ref["status"] == 0).all()
    assert np.array_equal(o["ticks"], ref["n_ticks"])
    assert np.array_equal(o["sim"][..., 0], ref["sim"][..., 0])
    np.testing.assert_array_equal(np.nan_to_num(o["sim"], nan=-1.0), np.nan_to_num(ref["sim"], nan=-1.0))


def test_mini_infer_surface_trains_a_toy_model():
    """pyro.infer / pyro.optim of the mini runtime (the surface of wn_network.py:285-302) on a
    conjugate-ish toy: every autoguide, looped and vectorised particles, scale handler."""
    from bayesair_b200.ppl import USING_REAL_PYRO, dist, poutine, pyro
    if USING_REAL_PYRO:
        pytest.skip("real pyro installed: its own infer module is used")
    y = torch.tensor([2.1, 1.9, 2.3, 2.0])

    def m(y):
        mu = pyro.sample("mu", dist.Normal(torch.tensor(0.0), torch.tensor(10.0)))
        s = pyro.sample("s", dist.Gamma(torch.tensor(2.0), torch.tensor(2.0)))
        pyro.factor("lik", dist.Normal(mu, s).log_prob(y.reshape((-1,) + (1,) * mu.dim())).sum(0))

    for name in ("AutoNormal", "AutoMultivariateNormal", "AutoDelta"):
        for vec in (True, False):
            pyro.clear_param_store()
            torch.manual_seed(0)
            model = poutine.scale(m, scale=0.5)
            guide = getattr(pyro.infer.autoguide, name)(model)
            svi = pyro.infer.SVI(model, guide, pyro.optim.ClippedAdam({"lr": 0.05, "lrd": 0.999}),
                                 pyro.infer.Trace_ELBO(num_particles=8, vectorize_particles=vec))
            losses = [svi.step(y) for _ in range(250)]
            assert np.mean(losses[-20:]) < np.mean(losses[:20]), (name, vec)
            med = guide.median(y)
            assert abs(float(med["mu"]) - 2.075) < 1.0, (name, vec, med)
            assert float(med["s"]) > 0


def test_trace_batches_sites_that_share_a_distribution():
    from bayesair_b200.ppl import USING_REAL_PYRO, dist, poutine, pyro
    if USING_REAL_PYRO:
        pytest.skip("mini runtime only")
    prior = dist.Gamma(torch.tensor(1.5), torch.tensor(10.0))

    def model():
        for i in range(40):
            pyro.sample(f"s{i}", prior)
        pyro.sample("odd", dist.Normal(torch.tensor(0.0), torch.tensor(1.0)))

    data = {f"s{i}": torch.rand(5) + 0.01 for i in range(40)}
    data["odd"] = torch.randn(5)
    tr = poutine.trace(poutine.scale(poutine.condition(model, data=data), scale=0.25)).get_trace()
    expect = 0.25 * (sum(prior.log_prob(data[f"s{i}"]).sum() for i in range(40))
                     + dist.Normal(torch.tensor(0.0), torch.tensor(1.0)).log_prob(data["odd"]).sum())
    assert abs(float(tr.log_prob_sum()) - float(expect)) < 1e-4
    assert tr.nodes["s3"]["log_prob"].shape == (5,)


def test_plan_cache_fingerprint_sees_edited_observations():
    """plan.states_fingerprint: the plan cache key changes when a flight's observations are
    replaced (obs stripped for a predictive run), edited in place, or flights are swapped at
    equal count -- and does not change otherwise."""
    import bayesair_b200 as ba
    from bayesair_b200.plan import states_fingerprint
    from bayesair_b200.synth import synth_day_frame
    fl, ap = ba.parse_schedule(synth_day_frame(4, 40))
    st = [ba.NetworkState({a.code: a for a in ap}, fl)]
    f0 = states_fingerprint(st)
    assert states_fingerprint(st) == f0
    keep = fl[3].actual_arrival_time
    fl[3].actual_arrival_time = None                       # stripped observation
    f1 = states_fingerprint(st)
    assert f1 != f0
    fl[3].actual_arrival_time = keep
    assert states_fingerprint(st) == f0
    fl[5].actual_departure_time.add_(0.25)                 # in-place edit
    f2 = states_fingerprint(st)
    assert f2 != f0
    st[0].pending_flights[7], st[0].pending_flights[8] = st[0].pending_flights[8], st[0].pending_flights[7]
    assert states_fingerprint(st) != f2


# ------------------------------ value mode: gradient w.r.t. the per-flight site values --
VALUE_GOLDENS = ["tiny4", "twoday5", "congested6", "saturated5", "obscancel4", "top10"]


def _value_golden(name):
    z = np.load(os.path.join(ROOT, "tests", "golden", name + "_valuegrad.npz"))
    return {k: z[k] for k in z.files}


def value_grad_error(got, ref, values):
    """Per-element relative error of a [D, F, 4] value gradient; flights without values
    (cancelled, padding) must come out exactly zero."""
    live = ~np.isnan(values)
    assert np.all(got[~live] == 0.0)
    scale = np.maximum(np.abs(ref[live]), 1e-4 * np.abs(ref[live]).max())
    return float((np.abs(got[live] - ref[live]) / scale).max())


@pytest.mark.parametrize("name", VALUE_GOLDENS)
def test_emulated_value_gradients_vs_reference(name):
    """d logp / d (per-flight site values) of the ring-less scan (host emulation of the kernel
    source) against autograd of the UNMODIFIED reference with those sites conditioned on
    leaf tensors (tests/golden/make_value_golden.py)."""
    from tests.hostemu import EmuPlan
    g, v = Golden(name), _value_golden(name)
    plan = EmuPlan(g.specs, len(g.global_codes), delta_t=g.delta_t)
    lat = {k: x[None] for k, x in g.latents.items()}
    vals = np.nan_to_num(v["values"], nan=1.0)[None]
    for shuffle in (0, 5):
        o = plan.forward(lat, g.scales, vals, value_mode=True, scan2=True, shuffle_seed=shuffle,
                         want_value_grad=True)
        assert (o["status"] == 0).all()
        assert abs(float(o["logp"].sum()) - float(v["total"])) <= 1e-4 * abs(float(v["total"]))
        assert value_grad_error(o["value_grad"][0], v["grad_values"], v["values"]) <= 1e-3
        # the global-latent gradients of the same (value-mode) run are pinned too
        assert rel_err(o["grads"]["mean_service"][0], v["grad_mean_service"], floor=1e-4) <= 1e-3
        assert rel_err(o["grads"]["mean_turnaround"][0], v["grad_mean_turnaround"], floor=1e-4) <= 1e-3


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver times next to the GPU arm): one JSON
    line, same metric / unit / config dict as the native arm would print, cpu_baseline and e2e
    blocks describing the run."""
    import json
    import subprocess
    import sys
    import bench
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                          "--workload", "cfg1", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-500:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == bench.METRIC and line["unit"] == bench.UNIT
    assert line["config"] == bench.workload_config("cfg1")
    assert line["higher_is_better"] is True and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": bench.UNIT, "h2d_bytes_per_step": 0,
                           "d2h_bytes_per_step": 0}
