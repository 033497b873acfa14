"""Generate the golden fixtures under ``tests/golden/`` by running the
UNMODIFIED reference (``/root/reference/bayes_air``) through
``oracle/reference_runner.py``.  Only runs in the build container (the reference
is not present on the GPU box); the resulting ``*.npz`` files are committed.

    python tests/golden/make_golden.py            # all cases (the N=100 one takes ~1 min)
    python tests/golden/make_golden.py tiny4      # one case

Every fixture holds the complete input (day tables as produced from the
reference's own ``parse_schedule`` + ``NetworkState`` sort, latents, per-flight
standard noise, the three scale values) and the reference outputs:
``total`` = ``trace.log_prob_sum()``, ``per_flight`` = the same restricted to the
``day*`` sites (what the CUDA factor replaces), per-site-class sums, the values of
the four latent per-flight sites, and torch-autograd gradients of ``per_flight``
w.r.t. every global latent and the three scales.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from bayesair_b200.compiler import compile_states  # noqa: E402
from bayesair_b200.synth import synth_day_frame, synth_latents  # noqa: E402
from oracle.reference_runner import build_reference_states, run_reference  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

# name: (N, [F per day], service_scale, include_cancellations, aircraft_mean,
#        cancel_frac, scales, delta_t, seed)
CASES = {
    "tiny4":      (4, [90], 0.03, False, 8.0, 0.0, (0.1, 0.1, 0.1), 0.2, 11),
    "top10":      (10, [480], 0.03, False, 8.0, 0.0, (0.1, 0.1, 0.1), 0.2, 12),
    "congested6": (6, [400], 0.15, False, 8.0, 0.0, (0.08, 0.12, 0.09), 0.2, 13),
    "saturated5": (5, [300], 0.35, False, 8.0, 0.0, (0.1, 0.1, 0.1), 0.2, 14),
    "twoday5":    (5, [120, 100], 0.04, False, 8.0, 0.0, (0.1, 0.11, 0.12), 0.2, 15),
    "dt01":       (4, [60], 0.03, False, 8.0, 0.0, (0.1, 0.1, 0.1), 0.1, 16),
    "obscancel4": (4, [100], 0.03, False, 8.0, 0.08, (0.1, 0.1, 0.1), 0.2, 17),
    "cancel4":    (4, [120], 0.03, True, 2.0, 0.10, (0.1, 0.12, 0.09), 0.2, 18),
    "cancel5":    (5, [200], 0.03, True, 7.0, 0.05, (0.1, 0.1, 0.1), 0.2, 19),
    "cancel2day": (5, [110, 130], 0.05, True, 4.0, 0.05, (0.09, 0.1, 0.11), 0.2, 20),
    "wn100":      (100, [3876], 0.02, False, 8.0, 0.0, (0.1, 0.1, 0.1), 0.2, 21),
}


# posterior-predictive cases (every observation stripped: obs=None everywhere)
# name: (N, [F per day], service_scale, include_cancellations, aircraft_mean, delta_t, seed)
PREDICTIVE = {
    "pred4c":  (4, [80], 0.03, True, 3.0, 0.2, 31),
    "pred5":   (5, [120, 90], 0.05, False, 8.0, 0.2, 32),
    "pred4s":  (4, [100], 0.04, True, 1.2, 0.2, 33),
}


def make_predictive(name):
    N, Fs, svc, canc, nac, dt, seed = PREDICTIVE[name]
    frames = [synth_day_frame(N, F, d, seed=seed) for d, F in enumerate(Fs)]
    lat = {k: v[0] for k, v in synth_latents(
        N, 1, seed=seed, service_scale=svc, include_cancellations=canc,
        n_aircraft_mean=nac).items()}
    D, Fmax = len(Fs), max(Fs)
    rng = np.random.default_rng([seed, 271828])
    noise = np.zeros((D, Fmax, 8), np.float32)
    noise[..., :4] = flight_noise(D, Fmax, seed)
    noise[..., 4] = rng.random((D, Fmax))
    noise[..., 5] = rng.standard_normal((D, Fmax))
    noise[..., 6] = rng.standard_normal((D, Fmax))
    ref = run_reference(frames, lat, noise, delta_t=dt, include_cancellations=canc,
                        want_grad=False, strip_observations=True)
    codes, specs = compile_states(build_reference_states(frames))
    out = {
        "n_days": D, "global_codes": np.array(codes), "delta_t": dt, "max_t": 30.0,
        "include_cancellations": int(canc), "scales": ref["scale_values"], "noise": noise,
        "ref_sim": ref["sim"], "ref_per_flight": ref["per_flight"],
    }
    for k, v in lat.items():
        out["lat_" + k] = v
    for d, s in enumerate(specs):
        out[f"day{d}_codes"] = np.array(s.codes)
        out[f"day{d}_names"] = np.array(s.flight_names)
        for f in ("airport_global", "sched_dep", "origin", "dest"):
            out[f"day{d}_{f}"] = getattr(s, f)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    n_c = int(np.nansum(ref["sim"][..., 0] != 0))
    print(f"{name}: predictive, {n_c} cancelled, per_flight={ref['per_flight']:.3f} -> "
          f"{os.path.getsize(path)} B")


def flight_noise(D, Fmax, seed):
    rng = np.random.default_rng([seed, 31337])
    z = np.empty((D, Fmax, 4), np.float32)
    z[..., 0] = rng.exponential(size=(D, Fmax))
    z[..., 1] = rng.standard_normal((D, Fmax))
    z[..., 2] = rng.exponential(size=(D, Fmax))
    z[..., 3] = rng.standard_normal((D, Fmax))
    return z


def make_case(name):
    N, Fs, svc, canc, nac, cfrac, scales, dt, seed = CASES[name]
    frames = [synth_day_frame(N, F, d, seed=seed, cancel_frac=cfrac) for d, F in enumerate(Fs)]
    lat = {k: v[0] for k, v in synth_latents(
        N, 1, seed=seed, service_scale=svc, include_cancellations=canc,
        n_aircraft_mean=nac).items()}
    D, Fmax = len(Fs), max(Fs)
    noise = flight_noise(D, Fmax, seed)
    ref = run_reference(frames, lat, noise, delta_t=dt, include_cancellations=canc,
                        scales=scales)
    codes, specs = compile_states(build_reference_states(frames))
    assert codes == ref["codes"]
    out = {
        "n_days": D, "global_codes": np.array(codes),
        "delta_t": dt, "max_t": 30.0, "include_cancellations": int(canc),
        "scales": np.array(scales, np.float64), "noise": noise,
        "ref_total": ref["total"], "ref_per_flight": ref["per_flight"],
        "ref_values": ref["values"], "ref_n_nodes": ref["n_nodes"],
        "ref_grad_scales": ref["grad_scales_per_flight"],
        "ref_times": np.array([ref["t_trace"], ref["t_log_prob_sum"], ref["t_backward"]]),
    }
    for k, v in lat.items():
        out["lat_" + k] = v
    for k, v in ref["class_sums"].items():
        out["ref_class_" + k] = v
    for k, v in ref["grad_per_flight"].items():
        out["ref_grad_" + k] = v
    for d, s in enumerate(specs):
        out[f"day{d}_codes"] = np.array(s.codes)
        out[f"day{d}_names"] = np.array(s.flight_names)
        for f in ("airport_global", "sched_dep", "act_dep", "act_arr", "cancel_obs",
                  "origin", "dest"):
            out[f"day{d}_{f}"] = getattr(s, f)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: per_flight={ref['per_flight']:.4f} total={ref['total']:.4f} "
          f"nodes={ref['n_nodes']} t=({ref['t_trace']:.2f},{ref['t_log_prob_sum']:.2f},"
          f"{ref['t_backward']:.2f})s -> {os.path.getsize(path)} B")


if __name__ == "__main__":
    names = sys.argv[1:] or (list(CASES) + list(PREDICTIVE))
    for n in names:
        make_predictive(n) if n in PREDICTIVE else make_case(n)
