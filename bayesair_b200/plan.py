"""``SimulationPlan``: compiled network days on one GPU + tensor-level entry points.

Host side of the C ABI.  PyTorch is plumbing here (device memory, streams); the
work is done by ``libbayesair_b200.so``.
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _capi
from .compiler import DaySpec, compile_states

DEFAULT_MAX_TICKS = 20000

STATUS_MESSAGES = {
    _capi.PD_TICK_CAP: "simulation did not complete within max_ticks",
    _capi.PD_OVERFLOW: "a shared-memory list overflowed and the exact re-run was disabled",
    _capi.PD_UNOBSERVED: "a flight has no observed departure/arrival/cancellation "
                         "(predictive mode is not handled by this entry point)",
    _capi.PD_INTERNAL: "internal list invariant violated",
}


class SimulationStatusError(RuntimeError):
    pass


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _resolve_device(device) -> torch.device:
    if not torch.cuda.is_available():
        raise _capi.BayesAirLibraryError(
            "bayesair_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    if device.type != "cuda":
        raise _capi.BayesAirLibraryError(
            f"device={device}: bayesair_b200 runs the simulation on the GPU only")
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    return device


class SimulationPlan:
    """Static tables of ``D`` network days resident on one GPU."""

    def __init__(self, specs: Sequence[DaySpec], global_codes: Sequence[str],
                 delta_t: float = 0.1, max_t: float = 30.0,
                 include_cancellations: bool = False, device=None,
                 max_ticks: int = DEFAULT_MAX_TICKS):
        self.device = _resolve_device(device)
        self.specs = list(specs)
        self.global_codes = list(global_codes)
        self.delta_t, self.max_t = float(delta_t), float(max_t)
        self.include_cancellations = bool(include_cancellations)
        self._h = _capi.PlanHandle(self.specs, len(self.global_codes), delta_t, max_t,
                                   include_cancellations, self.device.index, max_ticks)
        i = self._h.info
        self.n_days, self.n_airports, self.max_flights = i.n_days, i.n_airports, i.max_flights
        self.n_routes, self.n_airport_latents = i.n_routes, i.n_airport_latents
        self.grad_stride = i.grad_stride
        self.block_threads, self.smem_bytes, self.resident_ctas = (
            i.block_threads, i.smem_bytes, i.resident_ctas)
        self.route_origin = torch.from_numpy(self._h.route_origin.astype(np.int64)).to(self.device)
        self.route_dest = torch.from_numpy(self._h.route_dest.astype(np.int64)).to(self.device)
        self.n_flights_total = int(sum(s.n_flights for s in self.specs))
        # any flight with obs=None (not counting the measured times of cancelled flights)
        self.has_unobserved = any(
            bool(np.any(s.cancel_obs < 0)) or
            bool(np.any((np.isnan(s.act_dep) | np.isnan(s.act_arr)) & (s.cancel_obs != 1)))
            for s in self.specs)

    @classmethod
    def from_states(cls, states, delta_t=0.1, max_t=30.0, include_cancellations=False,
                    device=None, max_ticks=DEFAULT_MAX_TICKS) -> "SimulationPlan":
        codes, specs = compile_states(states)
        return cls(specs, codes, delta_t, max_t, include_cancellations, device, max_ticks)

    # ---- tensor-level calls ---------------------------------------------------
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _chk(self, t: torch.Tensor, shape: Tuple[int, ...], name: str) -> torch.Tensor:
        if t.device != self.device or t.dtype != torch.float32:
            t = t.to(device=self.device, dtype=torch.float32)
        if tuple(t.shape) != tuple(shape):
            raise ValueError(f"{name}: expected shape {tuple(shape)}, got {tuple(t.shape)}")
        return t.contiguous()

    def forward(self, lat_airport: torch.Tensor, lat_route: torch.Tensor, scales: torch.Tensor,
                noise: Optional[torch.Tensor] = None, *, seed: int = 0,
                noise_is_value: bool = False, want_grad: bool = False,
                force_exact_lists: bool = False, no_overflow_rerun: bool = False,
                want_ticks: bool = False, want_pop_ticks: bool = False,
                replay_waits: bool = False,
                force_ring_lists: bool = False, seed_tensor: Optional[torch.Tensor] = None,
                particle_offset: int = 0, day_offset: int = 0,
                want_value_grad: bool = False) -> Dict[str, torch.Tensor]:
        """``lat_airport [P,C,N]``, ``lat_route [P,R]``, ``scales [3]``,
        ``noise [P,D,Fmax,4]`` or None (Philox from ``seed``).  Asynchronous.
        ``seed_tensor``: one int64 on the device, read by the kernel when it starts (CUDA
        graphs replay with a fresh seed); ``particle_offset``: index of the first particle in
        the global batch (ranks sharing a seed then draw different per-flight noise);
        ``day_offset``: index of this plan's first day in the whole data set (a rank whose plan
        holds a subset of the days then draws what the full plan would);
        ``want_value_grad`` (with ``noise_is_value``): also returns ``value_tape [P,D,Fmax,4]``,
        the gradient of ``logp[p, d]`` w.r.t. the per-flight site values in ``noise``."""
        P = int(lat_airport.shape[0])
        D, N, Cn, R, F = (self.n_days, self.n_airports, self.n_airport_latents, self.n_routes,
                          self.max_flights)
        la = self._chk(lat_airport, (P, Cn, N), "lat_airport")
        lr = self._chk(lat_route, (P, R), "lat_route")
        sc = self._chk(scales, (3,), "scales")
        nz = None if noise is None else self._chk(noise, (P, D, F, 4), "noise")
        with torch.cuda.device(self.device):
            logp = torch.empty((P, D), dtype=torch.float32, device=self.device)
            status = torch.empty((P, D), dtype=torch.int32, device=self.device)
            ticks = torch.empty((P, D), dtype=torch.int32, device=self.device) if want_ticks else None
            tape = (torch.empty((P, D, self.grad_stride), dtype=torch.float32, device=self.device)
                    if want_grad else None)
            pop = (torch.empty((P, D, F, 2), dtype=torch.int32, device=self.device)
                   if want_pop_ticks else None)
            vtape = (torch.empty((P, D, F, 4), dtype=torch.float32, device=self.device)
                     if want_value_grad else None)
            opts = _capi.BaForwardOpts(int(noise_is_value), int(want_grad), int(force_exact_lists),
                                       int(no_overflow_rerun), int(replay_waits),
                                       int(force_ring_lists), int(seed) & (2**64 - 1),
                                       None if pop is None else pop.data_ptr(),
                                       None if seed_tensor is None else seed_tensor.data_ptr(),
                                       int(particle_offset) & 0xFFFFFFFF, int(day_offset) & 0xFFFFFFFF,
                                       None if vtape is None else vtape.data_ptr())
            with torch.cuda.nvtx.range("ba.forward"):
                self._h.forward(P, _ptr(la), _ptr(lr), _ptr(sc), _ptr(nz), opts, _ptr(logp),
                                _ptr(status), _ptr(ticks), _ptr(tape), self._stream())
        return {"logp": logp, "status": status, "ticks": ticks, "tape": tape, "pop_tick": pop,
                "value_tape": vtape, "_keep": (la, lr, sc, nz, seed_tensor)}

    def predict(self, lat_airport: torch.Tensor, lat_route: torch.Tensor, scales: torch.Tensor,
                noise: Optional[torch.Tensor] = None, obs_noise: Optional[torch.Tensor] = None,
                *, seed: int = 0, want_ticks: bool = False,
                want_pop_ticks: bool = False) -> Dict[str, torch.Tensor]:
        """Posterior-predictive forward simulation (``ba_predict``): flights without
        observations are sampled.  Returns ``sim [P, D, Fmax, 3]`` = (cancelled 0/1,
        simulated departure time, simulated arrival time; NaN where cancelled / padded)."""
        P = int(lat_airport.shape[0])
        D, N, Cn, R, F = (self.n_days, self.n_airports, self.n_airport_latents, self.n_routes,
                          self.max_flights)
        la = self._chk(lat_airport, (P, Cn, N), "lat_airport")
        lr = self._chk(lat_route, (P, R), "lat_route")
        sc = self._chk(scales, (3,), "scales")
        nz = None if noise is None else self._chk(noise, (P, D, F, 4), "noise")
        n2 = None if obs_noise is None else self._chk(obs_noise, (P, D, F, 4), "obs_noise")
        with torch.cuda.device(self.device):
            sim = torch.full((P, D, F, 3), float("nan"), dtype=torch.float32, device=self.device)
            status = torch.empty((P, D), dtype=torch.int32, device=self.device)
            ticks = torch.empty((P, D), dtype=torch.int32, device=self.device) if want_ticks else None
            pop = (torch.empty((P, D, F, 2), dtype=torch.int32, device=self.device)
                   if want_pop_ticks else None)
            opts = _capi.BaForwardOpts(0, 0, 1, 0, 1, 0, int(seed) & (2**64 - 1),
                                       None if pop is None else pop.data_ptr())
            self._h.predict(P, _ptr(la), _ptr(lr), _ptr(sc), _ptr(nz), _ptr(n2), opts, _ptr(sim),
                            _ptr(status), _ptr(ticks), self._stream())
        return {"sim": sim, "status": status, "ticks": ticks, "pop_tick": pop,
                "_keep": (la, lr, sc, nz, n2)}

    def backward(self, dlogp: torch.Tensor, tape: torch.Tensor):
        P = int(tape.shape[0])
        dl = self._chk(dlogp, (P, self.n_days), "dlogp")
        with torch.cuda.device(self.device):
            ga = torch.empty((P, self.n_airport_latents, self.n_airports), dtype=torch.float32,
                             device=self.device)
            gr = torch.empty((P, self.n_routes), dtype=torch.float32, device=self.device)
            gs = torch.empty((P, 3), dtype=torch.float32, device=self.device)
            with torch.cuda.nvtx.range("ba.backward"):
                self._h.backward(P, _ptr(dl), _ptr(tape), _ptr(ga), _ptr(gr), _ptr(gs), self._stream())
        return ga, gr, gs

    def backward_values(self, dlogp: torch.Tensor, value_tape: torch.Tensor) -> torch.Tensor:
        """Cotangent w.r.t. the per-flight site values of a value-mode forward
        (``ba_backward_values``): ``[P, D, Fmax, 4]``."""
        P = int(value_tape.shape[0])
        dl = self._chk(dlogp, (P, self.n_days), "dlogp")
        with torch.cuda.device(self.device):
            g = torch.empty_like(value_tape)
            self._h.backward_values(P, _ptr(dl), _ptr(value_tape), _ptr(g), self._stream())
        return g

    def generate_noise(self, n_particles: int, seed: int) -> torch.Tensor:
        with torch.cuda.device(self.device):
            nz = torch.empty((n_particles, self.n_days, self.max_flights, 4), dtype=torch.float32,
                             device=self.device)
            self._h.generate_noise(n_particles, seed, _ptr(nz), self._stream())
        return nz

    def logjoint_host(self, lat_airport: np.ndarray, lat_route: np.ndarray, scales: np.ndarray,
                      noise: Optional[np.ndarray] = None, *, seed: int = 0, want_grad: bool = True,
                      noise_is_value: bool = False, out: Optional[dict] = None) -> dict:
        """Host-buffer end-to-end call (``ba_logjoint_host``): numpy / pinned arrays in,
        numpy arrays out; copies, both kernels and the final synchronise included."""
        lat_airport = np.asarray(lat_airport)
        P = int(lat_airport.shape[0])
        D, N, Cn, R = self.n_days, self.n_airports, self.n_airport_latents, self.n_routes

        def host(a, shape, name, dtype=np.float32):
            # the C side reads / writes raw pointers: dtype, shape and contiguity are checked here
            if a is None:
                return None
            b = np.ascontiguousarray(a, dtype=dtype)
            if tuple(b.shape) != tuple(shape):
                raise ValueError(f"{name}: expected shape {tuple(shape)}, got {tuple(b.shape)}")
            return b

        def host_out(a, shape, name, dtype=np.float32):
            if a is None:
                return None
            if not (isinstance(a, np.ndarray) and a.dtype == dtype and a.flags.c_contiguous
                    and a.flags.writeable and tuple(a.shape) == tuple(shape)):
                raise ValueError(f"out[{name!r}]: needs a writable C-contiguous {np.dtype(dtype).name} "
                                 f"array of shape {tuple(shape)}")
            return a

        lat_airport = host(lat_airport, (P, Cn, N), "lat_airport")
        lat_route = host(lat_route, (P, R), "lat_route")
        scales = host(scales, (3,), "scales")
        noise = host(noise, (P, D, self.max_flights, 4), "noise")
        if noise_is_value and noise is None:
            raise ValueError("noise_is_value needs a noise array")
        if out is None:
            out = {
                "logp": np.empty((P, D), np.float32), "status": np.empty((P, D), np.uint32),
                "g_airport": np.empty((P, Cn, N), np.float32) if want_grad else None,
                "g_route": np.empty((P, R), np.float32) if want_grad else None,
                "g_scales": np.empty((P, 3), np.float32) if want_grad else None,
            }
        host_out(out["logp"], (P, D), "logp")
        host_out(out["status"], (P, D), "status", np.uint32)
        if want_grad:
            host_out(out["g_airport"], (P, Cn, N), "g_airport")
            host_out(out["g_route"], (P, R), "g_route")
            host_out(out["g_scales"], (P, 3), "g_scales")
        vp = lambda a: None if a is None else C.c_void_p(a.ctypes.data)  # noqa: E731
        opts = _capi.BaForwardOpts(int(noise_is_value), int(want_grad), 0, 0, 0, 0,
                                   int(seed) & (2**64 - 1), None)
        # the call shares the plan's scratch with forward(): order it after whatever the
        # caller's torch stream still has in flight on this plan
        torch.cuda.current_stream(self.device).synchronize()
        self._h.logjoint_host(P, vp(lat_airport), vp(lat_route), vp(scales), vp(noise), opts,
                              vp(out["logp"]), vp(out["status"]), vp(out["g_airport"]),
                              vp(out["g_route"]), vp(out["g_scales"]))
        return out

    # ---- helpers ---------------------------------------------------------------
    def pack_latents(self, mean_turnaround, mean_service, travel, log_initial_aircraft=None,
                     base_cancel_logprob=None):
        """``[P,N]`` rows + dense ``[P,N,N]`` travel times -> kernel layout
        (``[P,C,N]``, ``[P,R]``); differentiable torch ops."""
        rows = [mean_turnaround, mean_service]
        if self.include_cancellations:
            rows += [log_initial_aircraft, base_cancel_logprob]
        lat_airport = torch.stack(rows, dim=1)
        lat_route = travel[:, self.route_origin, self.route_dest]
        return lat_airport, lat_route

    @staticmethod
    def check_status(status: torch.Tensor):
        """Raises if any particle-day did not complete cleanly (synchronises)."""
        bad = int(torch.bitwise_or(status, 0).ne(0).sum())
        if bad:
            bits = int(np.bitwise_or.reduce(status.detach().cpu().numpy().ravel().astype(np.int64)))
            why = "; ".join(m for b, m in STATUS_MESSAGES.items() if bits & b)
            raise SimulationStatusError(f"{bad} particle-day(s) failed: {why}")

    def close(self):
        self._h.close()


# ---------------------------------------------------------------------------
# plan cache: callers of the model reuse the same ``states`` objects for
# thousands of calls (the reference deep-copies them every call, model.py:38)
# ---------------------------------------------------------------------------
_PLAN_CACHE: "OrderedDict[tuple, Tuple[SimulationPlan, list]]" = OrderedDict()
_PLAN_CACHE_SIZE = 16


def _tensor_tag(t):
    # identity + in-place version counter: a replaced attribute (``f.actual_arrival_time =
    # None`` to make a flight unobserved) and an in-place edit (``t.fill_(..)``) both change it
    return None if t is None else (id(t), getattr(t, "_version", 0))


def states_fingerprint(states) -> int:
    """Cheap content fingerprint of what the schedule compiler reads from ``states``: per
    flight its identity, endpoints and the identity / version of its observation tensors.
    O(flights) Python, no device or tensor reads."""
    acc = []
    for s in states:
        acc.append(tuple(s.airports.keys()))
        acc.append(tuple(
            (id(f), f.origin, f.destination, _tensor_tag(f.scheduled_departure_time),
             _tensor_tag(f.actual_departure_time), _tensor_tag(f.actual_arrival_time),
             _tensor_tag(f.actually_cancelled))
            for f in s.pending_flights))
    return hash(tuple(acc))


def plan_for_states(states, delta_t, max_t, include_cancellations, device) -> SimulationPlan:
    """Compiled plan of ``states``, cached.  The key holds a content fingerprint
    (``states_fingerprint``), so editing a flight's observations -- in place or by
    assignment -- or swapping flights yields a fresh plan instead of the stale one."""
    device = _resolve_device(device)
    key = (tuple(id(s) for s in states), states_fingerprint(states),
           float(delta_t), float(max_t), bool(include_cancellations), device.index)
    hit = _PLAN_CACHE.get(key)
    if hit is not None:
        _PLAN_CACHE.move_to_end(key)
        return hit[0]
    with torch.cuda.nvtx.range("ba.plan_build") if torch.cuda.is_available() else _no_range():
        plan = SimulationPlan.from_states(states, delta_t, max_t, include_cancellations, device)
    # strong refs to the states and their tensors keep the ids in the key valid
    keep = [list(states), [(f.scheduled_departure_time, f.actual_departure_time,
                            f.actual_arrival_time, f.actually_cancelled)
                           for s in states for f in s.pending_flights]]
    _PLAN_CACHE[key] = (plan, keep)
    while len(_PLAN_CACHE) > _PLAN_CACHE_SIZE:
        # drop the reference only: a live autograd graph may still hold the plan; its
        # device memory is released when the last reference goes (PlanHandle.__del__)
        _PLAN_CACHE.popitem(last=False)
    return plan


class _no_range:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def clear_plan_cache():
    """Forgets every cached plan (their memory is freed once nothing else refers to them)."""
    _PLAN_CACHE.clear()
