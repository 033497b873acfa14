"""Schedule compiler: ``list[NetworkState]`` -> flat per-day tables.

This is the step immediately in front of the kernel.  It
turns the reference's Python object graph into struct-of-arrays tables with the
exact orderings the reference's semantics depend on:

* flights of a day in **stable scheduled-departure order**, compared in fp32
  (``NetworkState.__post_init__``, ``bayes_air/network.py:30-32``);
* airports of a day in that day's ``state.airports`` **dict order**
  (``bayes_air/schedule.py:100-104`` -- first appearance, may differ per day);
* latent indexing in ``states[0].airports`` order (``bayes_air/model.py:58``).

Nothing here touches the simulation itself.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np


@dataclass
class DaySpec:
    """One network day as numpy SoA (host side)."""

    codes: List[str]              # day-local airport order
    airport_global: np.ndarray    # int32 [N]: day-local index -> global (latent) index
    sched_dep: np.ndarray         # float32 [F]
    act_dep: np.ndarray           # float32 [F], NaN = unobserved
    act_arr: np.ndarray           # float32 [F], NaN = unobserved
    cancel_obs: np.ndarray        # int8 [F]: 0 / 1, -1 = unobserved
    origin: np.ndarray            # int32 [F] day-local airport index
    dest: np.ndarray              # int32 [F]
    flight_names: Optional[List[str]] = None   # str(flight): site-name stems

    @property
    def n_airports(self) -> int:
        return len(self.codes)

    @property
    def n_flights(self) -> int:
        return int(self.sched_dep.shape[0])


def _opt_time(x) -> float:
    return math.nan if x is None else float(x)


def compile_states(states: Sequence) -> Tuple[List[str], List[DaySpec]]:
    """Flatten ``states`` (reference or ``bayesair_b200`` NetworkState objects).

    Returns ``(global_codes, day_specs)``.  Inputs are only read."""
    if len(states) == 0:
        raise ValueError("states must hold at least one NetworkState")
    global_codes = [str(c) for c in states[0].airports.keys()]
    gindex = {c: i for i, c in enumerate(global_codes)}
    specs = []
    for d, state in enumerate(states):
        codes = [str(c) for c in state.airports.keys()]
        if set(codes) != set(global_codes):
            raise ValueError(
                f"day {d} has a different airport set than day 0 "
                "(the model requires identical airports on all days, model.py:26-27)")
        lindex = {c: i for i, c in enumerate(codes)}
        flights = list(state.pending_flights)
        sched = np.array([float(f.scheduled_departure_time) for f in flights], np.float32)
        order = np.argsort(sched, kind="stable")      # no-op for a built NetworkState
        flights = [flights[i] for i in order]
        F = len(flights)
        spec = DaySpec(
            codes=codes,
            airport_global=np.array([gindex[c] for c in codes], np.int32),
            sched_dep=sched[order].astype(np.float32),
            act_dep=np.array([_opt_time(f.actual_departure_time) for f in flights], np.float32),
            act_arr=np.array([_opt_time(f.actual_arrival_time) for f in flights], np.float32),
            cancel_obs=np.array(
                [-1 if f.actually_cancelled is None else int(float(f.actually_cancelled) != 0.0)
                 for f in flights], np.int8),
            origin=np.array([lindex[f.origin] for f in flights], np.int32).reshape(F),
            dest=np.array([lindex[f.destination] for f in flights], np.int32).reshape(F),
            flight_names=[str(f) for f in flights],
        )
        if len(set(spec.flight_names)) != F:
            raise ValueError(f"day {d}: duplicate flight identifiers (site names would collide)")
        specs.append(spec)
    return global_codes, specs


def day_spec_from_frame(df, global_codes: Optional[List[str]] = None) -> DaySpec:
    """Vectorised equivalent of ``parse_schedule`` + ``NetworkState`` +
    ``compile_states`` for one dataframe in the dataloader schema (used for the
    large benchmark workloads, where building 10^5 ``Flight`` objects is pure
    overhead).  Cancelled rows lose their measured times, as in
    ``bayes_air/schedule.py:45-50``."""
    import pandas as pd

    codes = [str(c) for c in pd.unique(
        pd.concat([df["origin_airport"], df["destination_airport"]]))]
    if global_codes is None:
        global_codes = codes
    gindex = {c: i for i, c in enumerate(global_codes)}
    lindex = {c: i for i, c in enumerate(codes)}
    sched = df["scheduled_departure_time"].to_numpy(np.float64).astype(np.float32)
    order = np.argsort(sched, kind="stable")
    canc = df["cancelled"].to_numpy(bool)[order]
    adep = df["actual_departure_time"].to_numpy(np.float64).astype(np.float32)[order]
    aarr = df["actual_arrival_time"].to_numpy(np.float64).astype(np.float32)[order]
    adep[canc] = np.nan
    aarr[canc] = np.nan
    o = df["origin_airport"].map(lindex).to_numpy(np.int32)[order]
    dd = df["destination_airport"].map(lindex).to_numpy(np.int32)[order]
    names = [
        f"{n}_{a}_{b}" for n, a, b in zip(
            df["flight_number"].to_numpy()[order],
            df["origin_airport"].to_numpy()[order],
            df["destination_airport"].to_numpy()[order])
    ]
    return DaySpec(
        codes=codes,
        airport_global=np.array([gindex[c] for c in codes], np.int32),
        sched_dep=sched[order], act_dep=adep, act_arr=aarr,
        cancel_obs=canc.astype(np.int8), origin=o, dest=dd, flight_names=names,
    )


def used_routes(specs: Sequence[DaySpec], n_global: int) -> np.ndarray:
    """Sorted unique ``origin_global * N + dest_global`` over all days: the
    compact route table (dense ``[N,N]`` travel times are gathered down to this,
    SURVEY.md H5)."""
    keys = []
    for s in specs:
        g = s.airport_global.astype(np.int64)
        keys.append(g[s.origin] * n_global + g[s.dest])
    if not keys:
        return np.zeros(0, np.int64)
    return np.unique(np.concatenate(keys))
