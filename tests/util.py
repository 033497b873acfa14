"""Shared helpers for the tests: golden-fixture loading and noise tables."""
from __future__ import annotations

import glob
import os

import numpy as np

from bayesair_b200.compiler import DaySpec

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
LATENT_KEYS = ("mean_turnaround", "mean_service", "travel",
               "log_initial_aircraft", "base_cancel_logprob")


def _all_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def golden_names():
    """Fully observed cases (log-prob + gradient fixtures)."""
    return [n for n in _all_names() if not n.startswith("pred") and not n.endswith("_valuegrad")]


def predictive_names():
    """Posterior-predictive cases (every obs=None; simulated outcomes fixtures)."""
    return [n for n in _all_names() if n.startswith("pred")]


class PredictiveGolden:
    """One posterior-predictive fixture written by tests/golden/make_golden.py."""

    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
        self.name = name
        self.n_days = int(z["n_days"])
        self.global_codes = [str(c) for c in z["global_codes"]]
        self.delta_t, self.max_t = float(z["delta_t"]), float(z["max_t"])
        self.include_cancellations = bool(int(z["include_cancellations"]))
        self.scales = tuple(float(s) for s in z["scales"])
        self.noise = z["noise"][..., :4]            # [D, Fmax, 4]
        self.obs_noise = np.ascontiguousarray(z["noise"][..., 4:8])
        self.latents = {k: z["lat_" + k] for k in LATENT_KEYS if "lat_" + k in z}
        self.ref_sim = z["ref_sim"]                 # [D, Fmax, 3] cancelled, dep, arr
        self.specs = []
        for d in range(self.n_days):
            F = z[f"day{d}_sched_dep"].shape[0]
            nan = np.full(F, np.nan, np.float32)
            self.specs.append(DaySpec(
                codes=[str(c) for c in z[f"day{d}_codes"]],
                airport_global=z[f"day{d}_airport_global"],
                sched_dep=z[f"day{d}_sched_dep"], act_dep=nan.copy(), act_arr=nan.copy(),
                cancel_obs=np.full(F, -1, np.int8),
                origin=z[f"day{d}_origin"], dest=z[f"day{d}_dest"],
                flight_names=[str(c) for c in z[f"day{d}_names"]]))

    def check(self, sim, d):
        """sim: [F(max), 3] of day d from an implementation driven by the same noise."""
        F = self.specs[d].n_flights
        ref = self.ref_sim[d, :F]
        got = np.asarray(sim)[:F]
        cancelled = ref[:, 0] != 0
        assert np.array_equal(got[:, 0] != 0, cancelled)
        ok = ~cancelled
        np.testing.assert_allclose(got[ok, 1], ref[ok, 1], rtol=0, atol=4e-6)
        np.testing.assert_allclose(got[ok, 2], ref[ok, 2], rtol=0, atol=4e-6)


class Golden:
    """One fixture written by tests/golden/make_golden.py."""

    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
        self.name = name
        self.z = z
        self.n_days = int(z["n_days"])
        self.global_codes = [str(c) for c in z["global_codes"]]
        self.delta_t = float(z["delta_t"])
        self.max_t = float(z["max_t"])
        self.include_cancellations = bool(int(z["include_cancellations"]))
        self.scales = tuple(float(s) for s in z["scales"])
        self.noise = z["noise"]                       # [D, Fmax, 4]
        self.latents = {k: z["lat_" + k] for k in LATENT_KEYS if "lat_" + k in z}
        self.specs = []
        for d in range(self.n_days):
            self.specs.append(DaySpec(
                codes=[str(c) for c in z[f"day{d}_codes"]],
                airport_global=z[f"day{d}_airport_global"],
                sched_dep=z[f"day{d}_sched_dep"], act_dep=z[f"day{d}_act_dep"],
                act_arr=z[f"day{d}_act_arr"], cancel_obs=z[f"day{d}_cancel_obs"],
                origin=z[f"day{d}_origin"], dest=z[f"day{d}_dest"],
                flight_names=[str(c) for c in z[f"day{d}_names"]]))
        self.ref_per_flight = float(z["ref_per_flight"])
        self.ref_total = float(z["ref_total"])
        self.ref_values = z["ref_values"]             # [D, Fmax, 4]
        self.ref_class = {k[len("ref_class_"):]: float(z[k]) for k in z.files
                          if k.startswith("ref_class_")}
        self.ref_grads = {k[len("ref_grad_"):]: z[k] for k in z.files
                          if k.startswith("ref_grad_") and k != "ref_grad_scales"}
        self.ref_grad_scales = z["ref_grad_scales"]


def rel_err(a, b, floor=0.0):
    """Largest PER-ELEMENT relative error |a - b| / |b|.  Elements whose reference is
    exactly zero (unused routes, diagonal travel times) must be exactly zero too: there the
    absolute error is reported against the smallest normal number, i.e. any non-zero value
    fails every tolerance.  `floor` > 0 bounds the denominator of the NON-zero elements from
    below by floor * max|b|: an element many orders of magnitude below the largest of its
    family is a sum of cancelling fp32 terms, and the reference's own fp32 autograd carries
    the same absolute uncertainty there."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    if a.size == 0:
        return 0.0
    den = np.maximum(np.abs(b), 1e-30)
    if floor > 0.0:
        den = np.where(b != 0.0, np.maximum(den, floor * np.abs(b).max()), den)
    return float((np.abs(a - b) / den).max())


def max_norm_err(a, b):
    """|a - b|_inf / |b|_inf (for quantities that are sums with cancellation, e.g. log-probs)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def std_noise(P, D, Fmax, seed):
    """[P, D, Fmax, 4] standard draws: Exp(1), N(0,1), Exp(1), N(0,1)."""
    rng = np.random.default_rng([seed, 424242])
    z = np.empty((P, D, Fmax, 4), np.float32)
    z[..., 0] = rng.exponential(size=(P, D, Fmax))
    z[..., 1] = rng.standard_normal((P, D, Fmax))
    z[..., 2] = rng.exponential(size=(P, D, Fmax))
    z[..., 3] = rng.standard_normal((P, D, Fmax))
    return z
