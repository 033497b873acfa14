"""ORACLE / TEST INFRASTRUCTURE ONLY -- ctypes front end of ``oracle/ba_oracle.c``.

Builds ``oracle/_build/libba_oracle.so`` with gcc on first use (``build()`` is also
called from ``__graft_entry__.build()`` so the file travels to the GPU box) and
exposes ``run_day`` (one particle-day, with diagnostics) and ``run_batch``
(``P x D`` particle-days over a pthread pool; the CPU baseline of ``bench.py``).
Never imported by ``bayesair_b200``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "ba_oracle.c")
_OUT_DIR = os.path.join(_HERE, "_build")
_LIB_PATH = os.path.join(_OUT_DIR, "libba_oracle.so")
_lib = None

CLASS_NAMES = ("departure_service_time", "travel_time", "arrival_service_time",
               "turnaround_time", "cancelled", "simulated_departure_time",
               "simulated_arrival_time")


class _Day(C.Structure):
    _fields_ = [
        ("n_airports", C.c_int), ("n_flights", C.c_int),
        ("sched_dep", C.c_void_p), ("act_dep", C.c_void_p), ("act_arr", C.c_void_p),
        ("cancel_obs", C.c_void_p), ("origin", C.c_void_p), ("dest", C.c_void_p),
        ("airport_global", C.c_void_p),
    ]


class _Latents(C.Structure):
    _fields_ = [
        ("n_global", C.c_int),
        ("mean_turnaround", C.c_void_p), ("mean_service", C.c_void_p),
        ("travel", C.c_void_p), ("log_init_aircraft", C.c_void_p),
        ("base_cancel_logprob", C.c_void_p), ("scales", C.c_float * 3),
    ]


class _Result(C.Structure):
    _fields_ = [("logp", C.c_double), ("class_sums", C.c_double * 7),
                ("n_ticks", C.c_int), ("status", C.c_int)]


class _Grads(C.Structure):
    _fields_ = [("g_turn", C.c_void_p), ("g_serv", C.c_void_p), ("g_travel", C.c_void_p),
                ("g_logn0", C.c_void_p), ("g_basecl", C.c_void_p), ("g_scales", C.c_void_p)]


def build(force: bool = False) -> str:
    os.makedirs(_OUT_DIR, exist_ok=True)
    stale = (not os.path.exists(_LIB_PATH)
             or (os.path.exists(_SRC) and os.path.getmtime(_SRC) > os.path.getmtime(_LIB_PATH)))
    if force or stale:
        subprocess.check_call([
            "gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared",
            "-o", _LIB_PATH, _SRC, "-lm", "-lpthread"])
    return _LIB_PATH


def _load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.ba_oracle_run_day.restype = C.c_int
        _lib.ba_oracle_run_batch.restype = C.c_int
    return _lib


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class _DayHolder:
    """Keeps the contiguous arrays of a DaySpec alive next to the C struct."""

    def __init__(self, spec):
        self.arrs = dict(
            sched_dep=np.ascontiguousarray(spec.sched_dep, np.float32),
            act_dep=np.ascontiguousarray(spec.act_dep, np.float32),
            act_arr=np.ascontiguousarray(spec.act_arr, np.float32),
            cancel_obs=np.ascontiguousarray(spec.cancel_obs, np.int8),
            origin=np.ascontiguousarray(spec.origin, np.int32),
            dest=np.ascontiguousarray(spec.dest, np.int32),
            airport_global=np.ascontiguousarray(spec.airport_global, np.int32),
        )
        self.n_airports = spec.n_airports
        self.n_flights = spec.n_flights

    def fill(self, d: _Day):
        d.n_airports, d.n_flights = self.n_airports, self.n_flights
        for k, a in self.arrs.items():
            setattr(d, k, a.ctypes.data)


def run_day(spec, latents: Dict[str, np.ndarray], scales: Sequence[float],
            noise: np.ndarray, *, delta_t: float = 0.2, max_t: float = 30.0,
            include_cancellations: bool = False, noise_is_value: bool = False,
            want_grad: bool = True, max_ticks: int = 100000,
            obs_noise: Optional[np.ndarray] = None) -> Dict[str, object]:
    """One particle-day.  ``latents``: 1-particle arrays in global airport order
    (``mean_turnaround [Ng]``, ``mean_service [Ng]``, ``travel [Ng,Ng]``
    [+ ``log_initial_aircraft``, ``base_cancel_logprob``]); ``noise``: ``[F,4]``."""
    lib = _load()
    holder = _DayHolder(spec)
    day = _Day()
    holder.fill(day)
    F = spec.n_flights
    Ng = int(np.asarray(latents["mean_service"]).shape[0])
    keep = {k: np.ascontiguousarray(v, np.float32) for k, v in latents.items()}
    lat = _Latents()
    lat.n_global = Ng
    lat.mean_turnaround = keep["mean_turnaround"].ctypes.data
    lat.mean_service = keep["mean_service"].ctypes.data
    lat.travel = keep["travel"].ctypes.data
    if include_cancellations:
        lat.log_init_aircraft = keep["log_initial_aircraft"].ctypes.data
        lat.base_cancel_logprob = keep["base_cancel_logprob"].ctypes.data
    lat.scales = (C.c_float * 3)(*[float(s) for s in scales])
    nz = np.ascontiguousarray(noise[:F], np.float32)
    assert nz.shape == (F, 4)
    res = _Result()
    g = dict(turn=np.zeros(Ng), serv=np.zeros(Ng), travel=np.zeros((Ng, Ng)),
             logn0=np.zeros(Ng), basecl=np.zeros(Ng), scales=np.zeros(3))
    grads = _Grads(*[g[k].ctypes.data for k in
                     ("turn", "serv", "travel", "logn0", "basecl", "scales")])
    sim = np.zeros((F, 3), np.float32)
    pop = np.zeros((F, 2), np.int32)
    on = None if obs_noise is None else np.ascontiguousarray(obs_noise[:F], np.float32)
    assert on is None or on.shape == (F, 4)
    rc = lib.ba_oracle_run_day(
        C.byref(day), C.byref(lat), _ptr(nz), _ptr(on), C.c_int(int(noise_is_value)),
        C.c_float(delta_t), C.c_float(max_t), C.c_int(int(include_cancellations)),
        C.c_int(max_ticks), C.byref(res), C.byref(grads) if want_grad else None,
        _ptr(sim), _ptr(pop))
    out = {
        "status": rc, "logp": res.logp, "n_ticks": res.n_ticks,
        "class_sums": {n: res.class_sums[i] for i, n in enumerate(CLASS_NAMES)},
        "sim": sim, "pop_tick": pop,
    }
    if want_grad:
        out["grads"] = {
            "mean_turnaround": g["turn"], "mean_service": g["serv"], "travel": g["travel"],
            "log_initial_aircraft": g["logn0"], "base_cancel_logprob": g["basecl"],
            "scales": g["scales"],
        }
    return out


def run_batch(specs: Sequence, latents: Dict[str, np.ndarray], scales: Sequence[float],
              noise: np.ndarray, *, delta_t: float = 0.2, max_t: float = 30.0,
              include_cancellations: bool = False, noise_is_value: bool = False,
              want_grad: bool = True, n_threads: int = 1,
              max_ticks: int = 100000, obs_noise: Optional[np.ndarray] = None,
              want_sim: bool = False) -> Dict[str, np.ndarray]:
    """``P x D`` particle-days.  ``latents``: ``[P,Ng]`` / ``[P,Ng,Ng]`` arrays;
    ``noise``: ``[P, D, Fmax, 4]``.  Gradients are summed over days (unit
    cotangent on every ``logp[p,d]``)."""
    lib = _load()
    D = len(specs)
    holders = [_DayHolder(s) for s in specs]
    days = (_Day * D)()
    for h, d in zip(holders, days):
        h.fill(d)
    keep = {k: np.ascontiguousarray(v, np.float32) for k, v in latents.items()}
    P, Ng = keep["mean_service"].shape
    nz = np.ascontiguousarray(noise, np.float32)
    Fmax = nz.shape[2]
    assert nz.shape == (P, D, Fmax, 4) and Fmax >= max(s.n_flights for s in specs)
    sc = np.ascontiguousarray(np.asarray(scales, np.float32))
    on = None if obs_noise is None else np.ascontiguousarray(obs_noise, np.float32)
    assert on is None or on.shape == (P, D, Fmax, 4)
    sim = np.full((P, D, Fmax, 3), np.nan, np.float32) if want_sim else None
    logp = np.zeros((P, D)); status = np.zeros((P, D), np.int32)
    n_ticks = np.zeros((P, D), np.int32)
    g = dict(turn=np.zeros((P, Ng)), serv=np.zeros((P, Ng)), travel=np.zeros((P, Ng, Ng)),
             logn0=np.zeros((P, Ng)), basecl=np.zeros((P, Ng)), scales=np.zeros((P, 3)))
    lib.ba_oracle_run_batch(
        days, C.c_int(D), C.c_int(P), C.c_int(Ng), C.c_int(Fmax),
        _ptr(keep["mean_turnaround"]), _ptr(keep["mean_service"]), _ptr(keep["travel"]),
        _ptr(keep.get("log_initial_aircraft")) if include_cancellations else None,
        _ptr(keep.get("base_cancel_logprob")) if include_cancellations else None,
        _ptr(sc), _ptr(nz), _ptr(on), _ptr(sim), C.c_int(int(noise_is_value)),
        C.c_float(delta_t), C.c_float(max_t), C.c_int(int(include_cancellations)),
        C.c_int(max_ticks), C.c_int(int(want_grad)), C.c_int(int(n_threads)),
        _ptr(logp), _ptr(status), _ptr(n_ticks),
        _ptr(g["turn"]), _ptr(g["serv"]), _ptr(g["travel"]),
        _ptr(g["logn0"]), _ptr(g["basecl"]), _ptr(g["scales"]))
    out = {"logp": logp, "status": status, "n_ticks": n_ticks, "sim": sim}
    if want_grad:
        out["grads"] = {
            "mean_turnaround": g["turn"], "mean_service": g["serv"], "travel": g["travel"],
            "log_initial_aircraft": g["logn0"], "base_cancel_logprob": g["basecl"],
            "scales": g["scales"],
        }
    return out
