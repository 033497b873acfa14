"""Pins the C oracle (oracle/ba_oracle.c) against golden vectors produced by the
UNMODIFIED reference (tests/golden/make_golden.py): log_prob_sum of the
per-flight sites, every site-class partial sum, and autograd gradients."""
import numpy as np
import pytest

from oracle import ba_oracle
from tests.util import Golden, PredictiveGolden, golden_names, predictive_names, rel_err

LOGP_RTOL = 2e-6      # observed <= 6e-7; the north-star gate for the kernel is 1e-4
NORTH_STAR_LOGP_RTOL = 1e-4
GRAD_RTOL = 1e-4          # per element; the north star asks for 1e-3


def _run(g, noise_is_value=False):
    logp, classes = 0.0, {}
    grads = None
    for d, spec in enumerate(g.specs):
        nz = g.ref_values[d] if noise_is_value else g.noise[d]
        o = ba_oracle.run_day(spec, g.latents, g.scales, nz, delta_t=g.delta_t,
                              max_t=g.max_t, include_cancellations=g.include_cancellations,
                              noise_is_value=noise_is_value)
        assert o["status"] == 0
        logp += o["logp"]
        for k, v in o["class_sums"].items():
            classes[k] = classes.get(k, 0.0) + v
        if grads is None:
            grads = {k: np.array(v, np.float64) for k, v in o["grads"].items()}
        else:
            for k, v in o["grads"].items():
                grads[k] += v
    return logp, classes, grads


@pytest.mark.parametrize("name", golden_names())
def test_logp_and_site_classes(name):
    g = Golden(name)
    logp, classes, _ = _run(g)
    # (a) against the reference's per-site fp32 log-probs summed in double
    ref_sum = sum(g.ref_class.values())
    assert abs(logp - ref_sum) <= LOGP_RTOL * abs(ref_sum)
    # (b) against Trace.log_prob_sum() itself, which accumulates ~7F terms
    # sequentially in fp32: at F=3876 its own rounding error is
    # ~9e-5 relative, hence the north-star tolerance of 1e-4.
    assert abs(logp - g.ref_per_flight) <= NORTH_STAR_LOGP_RTOL * abs(g.ref_per_flight)
    for k, ref in g.ref_class.items():
        assert abs(classes[k] - ref) <= LOGP_RTOL * max(abs(ref), abs(g.ref_per_flight) * 1e-2), k


@pytest.mark.parametrize("name", golden_names())
def test_gradients(name):
    g = Golden(name)
    _, _, grads = _run(g)
    for k, ref in g.ref_grads.items():
        assert rel_err(grads[k], ref) <= GRAD_RTOL, k
    assert rel_err(grads["scales"], g.ref_grad_scales) <= GRAD_RTOL


@pytest.mark.parametrize("name", ["tiny4", "cancel4", "twoday5"])
def test_value_mode_logp(name):
    """Feeding the recorded site VALUES (x, tau, x, u) instead of standard noise
    must reproduce the same log-prob (the AutoMVN calling convention)."""
    g = Golden(name)
    logp, _, _ = _run(g, noise_is_value=True)
    ref_sum = sum(g.ref_class.values())
    assert abs(logp - ref_sum) <= LOGP_RTOL * abs(ref_sum)


@pytest.mark.parametrize("name", predictive_names())
def test_posterior_predictive_outcomes(name):
    """obs=None everywhere (posterior-predictive run): simulated
    cancellations identical, simulated departure/arrival times within 2 ulp of the
    reference driven by the same standard noise."""
    g = PredictiveGolden(name)
    for d, spec in enumerate(g.specs):
        o = ba_oracle.run_day(spec, g.latents, g.scales, g.noise[d], delta_t=g.delta_t,
                              max_t=g.max_t, include_cancellations=g.include_cancellations,
                              want_grad=False, obs_noise=g.obs_noise[d])
        assert o["status"] == 0
        g.check(o["sim"], d)
