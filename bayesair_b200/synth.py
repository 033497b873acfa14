"""Seeded synthetic WN-like network days.

The reference's flight data (``data/wn_dec21_dec30.csv`` …) is absent from the
mount (``/root/reference/.MISSING_LARGE_BLOBS``), so every test / benchmark
workload is a synthetic day with the dataframe schema the reference's
dataloader emits (``bayes_air/utils/dataloader.py:159-171``) and the size
statistics of the WN notebooks: hub-skewed origin /
destination choice, scheduled times on a one-minute grid, aircraft flown as
rotations (≈5–6 legs per aircraft per day, so that the cancellation mode's
aircraft bookkeeping sees a plausible fleet), ``actual_dep = sched + Exp(0.25 h)``
(never earlier than the aircraft's own previous arrival + 20 min) and
``actual_arr = actual_dep + route_time + N(0, 0.1 h)``.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import pandas as pd

WN_TOP10 = ["DEN", "DAL", "MDW", "PHX", "HOU", "LAS", "MCO", "BNA", "BWI", "OAK"]

COLUMNS = [
    "flight_number", "date", "origin_airport", "destination_airport",
    "scheduled_departure_time", "scheduled_arrival_time",
    "actual_departure_time", "actual_arrival_time",
    "wheels_off_time", "wheels_on_time", "cancelled",
]


def airport_codes(n: int) -> List[str]:
    """``n`` distinct three-letter codes; the ten busiest WN stations first."""
    codes = list(WN_TOP10[:n])
    i = 0
    while len(codes) < n:
        c = "Q" + chr(ord("A") + (i // 26) % 26) + chr(ord("A") + i % 26)
        if i >= 26 * 26:
            c = "Z%02X" % (i - 26 * 26) if i - 26 * 26 < 256 else "Y%03X" % i
        if c not in codes:
            codes.append(c)
        i += 1
    return codes


def route_times(n_airports: int, seed: int = 0) -> np.ndarray:
    """Nominal block time per ordered pair, hours in [1, 4]; fixed by ``seed``
    (shared by all days of a workload)."""
    rng = np.random.default_rng([seed, 7919])
    base = rng.uniform(1.0, 4.0, size=(n_airports, n_airports))
    base = 0.5 * (base + base.T)
    np.fill_diagonal(base, 0.0)
    return np.round(base * 60.0) / 60.0


def route_network(n_airports: int, seed: int = 0, zipf_s: float = 0.6,
                  routes_per_airport: float = 15.0) -> np.ndarray:
    """Boolean ``[N,N]`` adjacency of the served (directed, symmetric) routes.  The
    real WN network flies ~1 500 of the 9 900 possible pairs of its top-100
    airports; small networks are fully connected."""
    N = n_airports
    adj = ~np.eye(N, dtype=bool)
    target = int(routes_per_airport * N) // 2           # undirected pairs
    if N * (N - 1) // 2 <= target:
        return adj
    rng = np.random.default_rng([seed, 4242])
    w = 1.0 / np.arange(1, N + 1) ** zipf_s
    score = np.log(np.outer(w, w)) + rng.gumbel(size=(N, N)) * 4.0
    score = np.triu(score, 1) + np.tril(np.full((N, N), -np.inf))
    iu = np.dstack(np.unravel_index(np.argsort(-score, axis=None), (N, N)))[0][:target]
    adj = np.zeros((N, N), bool)
    adj[iu[:, 0], iu[:, 1]] = True
    adj |= adj.T
    for i in range(N):                                   # everyone reaches two hubs
        for h in (0, 1, 2):
            if h != i and adj[i].sum() < 2:
                adj[i, h] = adj[h, i] = True
    return adj


def synth_day_frame(
    n_airports: int,
    n_flights: int,
    day_index: int = 0,
    seed: int = 0,
    cancel_frac: float = 0.0,
    zipf_s: float = 0.6,
    date: Optional[str] = None,
) -> pd.DataFrame:
    """One synthetic network day as a dataframe (rows sorted by scheduled
    departure time, as ``dataloader.py:82-83`` leaves them)."""
    if n_flights < n_airports:
        raise ValueError("need at least one departure per airport")
    rng = np.random.default_rng([seed, day_index])
    codes = airport_codes(n_airports)
    w = 1.0 / np.arange(1, n_airports + 1) ** zipf_s
    w /= w.sum()
    rt = route_times(n_airports, seed)
    adj = route_network(n_airports, seed, zipf_s)
    dest_p = [np.where(adj[i], w, 0.0) / np.where(adj[i], w, 0.0).sum() for i in range(n_airports)]

    orig = np.empty(n_flights, np.int64)
    dest = np.empty(n_flights, np.int64)
    sdep = np.empty(n_flights)
    sarr = np.empty(n_flights)
    adep = np.empty(n_flights)
    aarr = np.empty(n_flights)

    def pick_other(cur):
        return int(rng.choice(n_airports, p=dest_p[cur]))

    # every airport must appear (the model requires identical airport sets on
    # all days): aircraft i < n_airports starts at airport i.
    k = 0
    tail = 0
    while k < n_flights:
        cur = tail if tail < n_airports else int(rng.choice(n_airports, p=w))
        tail += 1
        t_sched = float(rng.integers(5 * 60, 11 * 60)) / 60.0
        t_free = 0.0  # when the airframe is really available
        legs = int(rng.integers(4, 8))
        for _ in range(legs):
            if k >= n_flights or t_sched > 22.5:
                break
            d = pick_other(cur)
            block = rt[cur, d]
            orig[k], dest[k] = cur, d
            sdep[k] = t_sched
            sarr[k] = t_sched + block
            dep = max(t_sched + rng.exponential(0.25), t_free)
            arr = dep + max(0.3, block + rng.normal(0.0, 0.1))
            adep[k], aarr[k] = dep, arr
            t_free = arr + 1.0 / 3.0
            turn = float(rng.integers(35, 75)) / 60.0
            t_sched = round((sarr[k] + turn) * 60.0) / 60.0
            cur = d
            k += 1
    if len(set(orig.tolist()) | set(dest.tolist())) != n_airports:
        raise ValueError(
            "n_flights too small for every airport to appear "
            f"(N={n_airports}, F={n_flights}); use F >= 7*N")

    cancelled = rng.random(n_flights) < cancel_frac
    order = np.argsort(sdep, kind="stable")
    date = date or "2022-12-%02d" % (21 + day_index)
    df = pd.DataFrame({
        "flight_number": (1000 + np.arange(n_flights))[order],
        "date": [date] * n_flights,
        "origin_airport": [codes[i] for i in orig[order]],
        "destination_airport": [codes[i] for i in dest[order]],
        "scheduled_departure_time": sdep[order],
        "scheduled_arrival_time": sarr[order],
        "actual_departure_time": adep[order],
        "actual_arrival_time": aarr[order],
        "wheels_off_time": adep[order] + 0.2,
        "wheels_on_time": aarr[order] - 0.1,
        "cancelled": cancelled[order],
    }, columns=COLUMNS)
    return df.reset_index(drop=True)


def synth_latents(
    n_airports: int,
    n_particles: int,
    seed: int = 0,
    service_scale: float = 0.03,
    spread: float = 0.3,
    include_cancellations: bool = False,
    n_aircraft_mean: float = 8.0,
) -> Dict[str, np.ndarray]:
    """Benchmark / test latents in the ranges of the reference's posterior
    samples (``analysis/top_3_posterior_samples.csv``).
    Returns fp32 arrays ``mean_turnaround [P,N]``, ``mean_service [P,N]``,
    ``travel [P,N,N]`` (+ ``log_initial_aircraft``, ``base_cancel_logprob``)."""
    rng = np.random.default_rng([seed, 104729])
    P, N = n_particles, n_airports
    rt = route_times(N, seed)
    out = {
        "mean_turnaround": 0.5 * np.exp(spread * rng.standard_normal((P, N))),
        "mean_service": service_scale * np.exp(spread * rng.standard_normal((P, N))),
        "travel": np.maximum(rt, 0.5)[None] * np.exp(0.1 * rng.standard_normal((P, N, N))),
    }
    if include_cancellations:
        out["log_initial_aircraft"] = (
            np.log(n_aircraft_mean) + 0.3 * rng.standard_normal((P, N)))
        out["base_cancel_logprob"] = -3.0 + 0.3 * rng.standard_normal((P, N))
    return {k: v.astype(np.float32) for k, v in out.items()}
