"""ORACLE / TEST INFRASTRUCTURE ONLY -- runs the UNMODIFIED reference model.

Imports ``bayes_air`` straight from ``/root/reference`` (never copied), with the
Pyro-surface shim in ``oracle/pyro_shim`` standing in for the absent
``pyro-ppl`` (real Pyro is used instead if it is ever importable), and drives
``bayes_air.model.air_traffic_network_model`` exactly the way the reference's
WN driver does (``scripts/wn/train.py:301-316``):

    trace(condition(model, data=latents)).get_trace(states=..., delta_t=...,
          device=cpu, include_cancellations=...)  ->  trace.log_prob_sum()
          -> backward()

It only works inside the build container (``/root/reference`` does not exist on
the GPU box); its outputs are committed as fixtures by
``tests/golden/make_golden.py``.
"""
from __future__ import annotations

import importlib
import os
import sys
import time
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

REFERENCE_ROOT = os.environ.get("BAYESAIR_REFERENCE_ROOT", "/root/reference")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pyro_shim")

SITE_KINDS = ("departure_service_time", "travel_time", "arrival_service_time",
              "turnaround_time")
OBS_KINDS = ("cancelled", "simulated_departure_time", "simulated_arrival_time")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "bayes_air"))


def _import_reference():
    """Returns (pyro, bayes_air.model, bayes_air.network, bayes_air.schedule)."""
    if not reference_available():
        raise RuntimeError(f"reference not mounted at {REFERENCE_ROOT}")
    try:
        import pyro  # noqa: F401  real Pyro, if it exists
    except ImportError:
        if _SHIM not in sys.path:
            sys.path.insert(0, _SHIM)
    if REFERENCE_ROOT not in sys.path:
        sys.path.append(REFERENCE_ROOT)
    pyro = importlib.import_module("pyro")
    model = importlib.import_module("bayes_air.model")
    network = importlib.import_module("bayes_air.network")
    schedule = importlib.import_module("bayes_air.schedule")
    return pyro, model, network, schedule


def build_reference_states(frames):
    """``scripts/wn/train.py:187-196``: one NetworkState per day frame."""
    _, _, network, schedule = _import_reference()
    states = []
    for df in frames:
        flights, airports = schedule.parse_schedule(df)
        states.append(network.NetworkState(
            airports={a.code: a for a in airports}, pending_flights=flights))
    return states


def run_reference(
    frames: Sequence,
    latents: Dict[str, np.ndarray],
    noise: Optional[np.ndarray],
    delta_t: float = 0.2,
    max_t: float = 30.0,
    include_cancellations: bool = False,
    scales: Sequence[float] = (0.1, 0.1, 0.1),
    want_grad: bool = True,
    strip_observations: bool = False,
    values: Optional[np.ndarray] = None,
) -> Dict[str, object]:
    """One particle through the reference.

    latents: 1-particle arrays ``mean_turnaround [N]``, ``mean_service [N]``,
        ``travel [N,N]`` (+ ``log_initial_aircraft [N]``, ``base_cancel_logprob [N]``),
        indexed in ``states[0].airports`` order.
    noise:   ``[D, Fmax, 4]`` standard draws (Exp(1), N(0,1), Exp(1), N(0,1)) per
        flight in *sorted pending order*; None -> the global torch RNG is used.  With
        ``strip_observations`` (posterior-predictive run: every ``obs=None``) pass
        ``[D, Fmax, 8]``: columns 4-6 drive the three otherwise-observed sites
        (U(0,1) for ``_cancelled``, N(0,1) for ``_simulated_departure_time`` and
        ``_simulated_arrival_time``).
    values:  ``[D, Fmax, 4]`` -- "value mode": the four latent per-flight sites
        (departure service time, travel time, arrival service time, turnaround time) are
        CONDITIONED on these numbers as autograd leaves instead of being drawn, which is what
        a guide over the per-flight sites does (``wn_network.py:285-302``,
        AutoMultivariateNormal); the output then also holds ``grad_values [D, Fmax, 4]``, the
        gradient of the per-flight log-probability w.r.t. them.
    Returns a dict of numpy values (see keys at the bottom).
    """
    pyro, model, _, _ = _import_reference()
    states = build_reference_states(frames)
    if strip_observations:
        for st in states:
            for fl in st.pending_flights:
                fl.actual_departure_time = None
                fl.actual_arrival_time = None
                fl.actually_cancelled = None
    codes = list(states[0].airports.keys())
    N = len(codes)
    flight_index = [
        {str(fl): i for i, fl in enumerate(st.pending_flights)} for st in states
    ]

    # ---- leaves ---------------------------------------------------------
    def leaf(a):
        return torch.tensor(np.asarray(a, dtype=np.float32), requires_grad=want_grad)

    lv = {k: leaf(v) for k, v in latents.items()}
    data = {}
    for i, c in enumerate(codes):
        data[f"{c}_mean_turnaround_time"] = lv["mean_turnaround"][i]
        data[f"{c}_mean_service_time"] = lv["mean_service"][i]
        if include_cancellations:
            data[f"{c}_log_initial_available_aircraft"] = lv["log_initial_aircraft"][i]
            data[f"{c}_base_cancel_logprob"] = lv["base_cancel_logprob"][i]
        for j, c2 in enumerate(codes):
            if i != j:
                data[f"travel_time_{c}_{c2}"] = lv["travel"][i, j]

    value_leaves = None
    if values is not None:
        value_leaves = torch.tensor(np.nan_to_num(np.asarray(values, np.float32), nan=1.0),
                                    requires_grad=want_grad)
        for d, ix in enumerate(flight_index):
            for stem, i in ix.items():
                for col, suffix in enumerate(SITE_KINDS):
                    data[f"day{d}_{stem}_{suffix}"] = value_leaves[d, i, col]

    pyro.clear_param_store()
    names = ("runway_use_time_std_dev", "travel_time_variation",
             "turnaround_time_variation")
    for n, v in zip(names, scales):
        pyro.param(n, torch.tensor(float(v)), constraint=pyro.distributions.constraints.positive)

    # ---- noise injection ------------------------------------------------
    def source(site_name: str, kind: str):
        if noise is None or not site_name.startswith("day"):
            return None
        day_str, rest = site_name[3:].split("_", 1)
        d = int(day_str)
        for col, suffix in enumerate(SITE_KINDS):
            if rest.endswith("_" + suffix):
                stem = rest[: -len(suffix) - 1]
                return float(noise[d, flight_index[d][stem], col])
        if noise.shape[-1] >= 7:
            for col, suffix in enumerate(OBS_KINDS):
                if rest.endswith("_" + suffix):
                    stem = rest[: -len(suffix) - 1]
                    return float(noise[d, flight_index[d][stem], 4 + col])
        return None

    if hasattr(pyro, "set_noise_source"):
        pyro.set_noise_source(source if noise is not None else None)
    elif noise is not None:
        raise RuntimeError("noise injection needs the oracle shim")

    t0 = time.perf_counter()
    try:
        tr = pyro.poutine.trace(
            pyro.poutine.condition(model.air_traffic_network_model, data=data)
        ).get_trace(states=states, delta_t=delta_t, max_t=max_t,
                    device=torch.device("cpu"),
                    include_cancellations=include_cancellations)
    finally:
        if hasattr(pyro, "set_noise_source"):
            pyro.set_noise_source(None)
    t_trace = time.perf_counter() - t0

    t0 = time.perf_counter()
    total = tr.log_prob_sum()
    t_lp = time.perf_counter() - t0
    per_flight = tr.log_prob_sum(lambda name, site: name.startswith("day"))

    D = len(states)
    Fmax = max(len(ix) for ix in flight_index)
    values = np.full((D, Fmax, 4), np.nan, np.float32)
    sim = np.full((D, Fmax, 3), np.nan, np.float32)  # cancelled, dep, arr
    class_sums = {k: 0.0 for k in SITE_KINDS + OBS_KINDS}
    for name, site in tr.nodes.items():
        if site["type"] != "sample" or not name.startswith("day"):
            continue
        day_str, rest = name[3:].split("_", 1)
        d = int(day_str)
        for col, suffix in enumerate(SITE_KINDS):
            if rest.endswith("_" + suffix):
                stem = rest[: -len(suffix) - 1]
                values[d, flight_index[d][stem], col] = float(site["value"].reshape(-1)[0])
                class_sums[suffix] += float(site["fn"].log_prob(site["value"]).sum())
                break
        else:
            for col, suffix in enumerate(OBS_KINDS):
                if rest.endswith("_" + suffix):
                    stem = rest[: -len(suffix) - 1]
                    sim[d, flight_index[d][stem], col] = float(site["value"].reshape(-1)[0])
                    class_sums[suffix] += float(site["fn"].log_prob(site["value"]).sum())
                    break

    store = pyro.get_param_store()
    out = {
        "scale_values": np.array([float(store.get(n)) if hasattr(store, "get") else float(pyro.param(n))
                                  for n in names], np.float64),
        "codes": codes,
        "total": float(total),
        "per_flight": float(per_flight),
        "class_sums": class_sums,
        "values": values,
        "sim": sim,
        "n_nodes": len(tr.nodes),
        "t_trace": t_trace,
        "t_log_prob_sum": t_lp,
    }

    if want_grad:
        store = pyro.get_param_store()
        raws = [dict(store.named_parameters())[n] for n in names]
        leaves = list(lv.values()) + raws
        t0 = time.perf_counter()
        g_pf = torch.autograd.grad(per_flight, leaves, retain_graph=True, allow_unused=True)
        out["t_backward"] = time.perf_counter() - t0
        g_tot = torch.autograd.grad(total, leaves, allow_unused=True, retain_graph=True)

        def npz(g, like):
            return (np.zeros(like.shape, np.float32) if g is None
                    else g.detach().numpy().astype(np.float32))

        keys = list(lv.keys())
        out["grad_per_flight"] = {k: npz(g, lv[k]) for k, g in zip(keys, g_pf)}
        out["grad_total"] = {k: npz(g, lv[k]) for k, g in zip(keys, g_tot)}
        # params are stored unconstrained (log); d/d(value) = d/d(raw) / value
        vals = [float(torch.exp(r)) for r in raws]
        out["grad_scales_per_flight"] = np.array(
            [0.0 if g is None else float(g) / v for g, v in zip(g_pf[len(keys):], vals)],
            np.float64)
        if value_leaves is not None:
            (gv,) = torch.autograd.grad(per_flight, [value_leaves], allow_unused=True)
            out["grad_values"] = np.zeros(value_leaves.shape, np.float32) if gv is None \
                else gv.detach().numpy().astype(np.float32)
    return out
