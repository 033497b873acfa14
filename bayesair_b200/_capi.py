"""ctypes binding of ``libbayesair_b200.so`` (C ABI: ``include/bayesair_b200.h``).

The CUDA library is the only implementation of the hot path.  If it is missing,
cannot be loaded, or no CUDA device is present, every entry point raises -- there
is no CPU fallback by design.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

from . import build as _build

_lib = None

# status bits of d_status (include/bayesair_b200.h)
PD_TICK_CAP, PD_OVERFLOW, PD_UNOBSERVED, PD_INTERNAL = 1, 2, 4, 8

EXPORTS = (
    "ba_version", "ba_last_error", "ba_plan_create", "ba_plan_destroy", "ba_plan_info",
    "ba_plan_routes", "ba_tape_bytes", "ba_forward", "ba_backward", "ba_predict",
    "ba_generate_noise",
    "ba_logjoint_host", "ba_launch_count",
    "ba_mf_work_bytes", "ba_mf_sample", "ba_mf_backward",
    "ba_rqs_forward", "ba_rqs_backward", "ba_backward_values",
)


class BaDayTable(C.Structure):
    _fields_ = [("n_airports", C.c_int32), ("n_flights", C.c_int32),
                ("h_sched_dep", C.c_void_p), ("h_act_dep", C.c_void_p),
                ("h_act_arr", C.c_void_p), ("h_cancel_obs", C.c_void_p),
                ("h_origin", C.c_void_p), ("h_dest", C.c_void_p),
                ("h_airport_global", C.c_void_p)]


class BaPlanInfo(C.Structure):
    _fields_ = [("n_days", C.c_int32), ("n_airports", C.c_int32), ("max_flights", C.c_int32),
                ("n_routes", C.c_int32), ("n_airport_latents", C.c_int32),
                ("grad_stride", C.c_int32), ("include_cancellations", C.c_int32),
                ("device", C.c_int32), ("delta_t", C.c_float), ("max_t", C.c_float),
                ("max_ticks", C.c_int32), ("block_threads", C.c_int32),
                ("smem_bytes", C.c_int32), ("resident_ctas", C.c_int32)]


class BaForwardOpts(C.Structure):
    _fields_ = [("noise_is_value", C.c_int32), ("want_grad", C.c_int32),
                ("force_exact_lists", C.c_int32), ("no_overflow_rerun", C.c_int32),
                ("replay_waits", C.c_int32), ("force_ring_lists", C.c_int32),
                ("seed", C.c_uint64), ("d_pop_tick", C.c_void_p),
                ("d_seed", C.c_void_p), ("particle_offset", C.c_uint32),
                ("day_offset", C.c_uint32), ("d_value_tape", C.c_void_p)]


class BayesAirLibraryError(RuntimeError):
    pass


def library_path() -> str:
    return _build.LIB_PATH


def load_library(build_if_missing: bool = True):
    """Loads (building it first if the sources are newer) the CUDA library."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB_PATH
    if build_if_missing and _build.needs_build():
        try:
            _build.build_library()
        except Exception as exc:  # nvcc absent on a box that only has the prebuilt .so
            if not os.path.exists(path):
                raise BayesAirLibraryError(
                    f"libbayesair_b200.so is not built and cannot be built here: {exc}") from exc
    if not os.path.exists(path):
        raise BayesAirLibraryError(
            f"{path} not found: run `python -c 'import __graft_entry__ as g; g.build()'`; "
            "bayesair_b200 has no CPU fallback")
    lib = C.CDLL(path)
    lib.ba_version.restype = C.c_int
    lib.ba_last_error.restype = C.c_char_p
    lib.ba_tape_bytes.restype = C.c_size_t
    lib.ba_tape_bytes.argtypes = [C.c_void_p, C.c_int32]
    lib.ba_launch_count.restype = C.c_uint64
    lib.ba_plan_destroy.restype = None
    lib.ba_plan_destroy.argtypes = [C.c_void_p]
    for name in ("ba_plan_create", "ba_plan_info", "ba_plan_routes", "ba_forward",
                 "ba_backward", "ba_predict", "ba_generate_noise", "ba_logjoint_host"):
        getattr(lib, name).restype = C.c_int
    lib.ba_plan_create.argtypes = [C.POINTER(BaDayTable), C.c_int32, C.c_int32, C.c_float,
                                   C.c_float, C.c_int32, C.c_int32, C.c_int32,
                                   C.POINTER(C.c_void_p)]
    lib.ba_plan_info.argtypes = [C.c_void_p, C.POINTER(BaPlanInfo)]
    lib.ba_plan_routes.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ba_forward.argtypes = [C.c_void_p, C.c_int32] + [C.c_void_p] * 4 + \
        [C.POINTER(BaForwardOpts)] + [C.c_void_p] * 5
    lib.ba_backward.argtypes = [C.c_void_p, C.c_int32] + [C.c_void_p] * 6
    lib.ba_predict.argtypes = [C.c_void_p, C.c_int32] + [C.c_void_p] * 5 + \
        [C.POINTER(BaForwardOpts)] + [C.c_void_p] * 4
    lib.ba_generate_noise.argtypes = [C.c_void_p, C.c_int32, C.c_uint64, C.c_void_p, C.c_void_p]
    lib.ba_logjoint_host.argtypes = [C.c_void_p, C.c_int32] + [C.c_void_p] * 4 + \
        [C.POINTER(BaForwardOpts)] + [C.c_void_p] * 5
    lib.ba_mf_work_bytes.restype = C.c_size_t
    lib.ba_mf_work_bytes.argtypes = [C.c_void_p, C.c_int32]
    lib.ba_mf_sample.restype = C.c_int
    lib.ba_mf_sample.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_uint64, C.c_void_p,
                                 C.c_uint32] + [C.c_void_p] * 5
    lib.ba_mf_backward.restype = C.c_int
    lib.ba_mf_backward.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_uint64, C.c_void_p,
                                   C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float,
                                   C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ba_backward_values.restype = C.c_int
    lib.ba_backward_values.argtypes = [C.c_void_p, C.c_int32] + [C.c_void_p] * 4
    lib.ba_rqs_forward.restype = C.c_int
    lib.ba_rqs_forward.argtypes = [C.c_void_p] * 4 + [C.c_int64, C.c_int32, C.c_float, C.c_void_p]
    lib.ba_rqs_backward.restype = C.c_int
    lib.ba_rqs_backward.argtypes = [C.c_void_p] * 6 + [C.c_int64, C.c_int32, C.c_float, C.c_void_p]
    _lib = lib
    return lib


def _check(rc: int, what: str):
    if rc != 0:
        msg = load_library().ba_last_error()
        raise BayesAirLibraryError(f"{what} failed (status {rc}): {msg.decode() if msg else ''}")


def launch_count() -> int:
    return int(load_library().ba_launch_count())


def make_day_tables(specs: Sequence):
    """DaySpec list -> (ctypes array of BaDayTable, keep-alive list)."""
    arr = (BaDayTable * len(specs))()
    keep = []
    for s, t in zip(specs, arr):
        a = [np.ascontiguousarray(s.sched_dep, np.float32),
             np.ascontiguousarray(s.act_dep, np.float32),
             np.ascontiguousarray(s.act_arr, np.float32),
             np.ascontiguousarray(s.cancel_obs, np.int8),
             np.ascontiguousarray(s.origin, np.int32),
             np.ascontiguousarray(s.dest, np.int32),
             np.ascontiguousarray(s.airport_global, np.int32)]
        keep.append(a)
        t.n_airports, t.n_flights = s.n_airports, s.n_flights
        (t.h_sched_dep, t.h_act_dep, t.h_act_arr, t.h_cancel_obs, t.h_origin, t.h_dest,
         t.h_airport_global) = [x.ctypes.data for x in a]
    return arr, keep


class PlanHandle:
    """Owns one ``BaPlan*``."""

    def __init__(self, specs: Sequence, n_airports: int, delta_t: float, max_t: float,
                 include_cancellations: bool, device_index: int, max_ticks: int):
        lib = load_library()
        arr, keep = make_day_tables(specs)
        h = C.c_void_p()
        _check(lib.ba_plan_create(arr, len(specs), n_airports, float(delta_t), float(max_t),
                                  int(bool(include_cancellations)), int(max_ticks),
                                  int(device_index), C.byref(h)), "ba_plan_create")
        del keep
        self._h = h
        info = BaPlanInfo()
        _check(lib.ba_plan_info(h, C.byref(info)), "ba_plan_info")
        self.info = info
        R = info.n_routes
        self.route_origin = np.zeros(R, np.int32)
        self.route_dest = np.zeros(R, np.int32)
        _check(lib.ba_plan_routes(h, self.route_origin.ctypes.data, self.route_dest.ctypes.data),
               "ba_plan_routes")

    @property
    def handle(self):
        if self._h is None:
            raise BayesAirLibraryError("plan already destroyed")
        return self._h

    def tape_bytes(self, P: int) -> int:
        return int(load_library().ba_tape_bytes(self.handle, int(P)))

    def forward(self, P, lat_airport, lat_route, scales, noise, opts: BaForwardOpts,
                logp, status, ticks, tape, stream):
        _check(load_library().ba_forward(self.handle, int(P), lat_airport, lat_route, scales,
                                         noise, C.byref(opts), logp, status, ticks, tape,
                                         stream), "ba_forward")

    def predict(self, P, lat_airport, lat_route, scales, noise, noise2, opts: BaForwardOpts,
                sim, status, ticks, stream):
        _check(load_library().ba_predict(self.handle, int(P), lat_airport, lat_route, scales,
                                         noise, noise2, C.byref(opts), sim, status, ticks,
                                         stream), "ba_predict")

    def backward(self, P, dlogp, tape, g_airport, g_route, g_scales, stream):
        _check(load_library().ba_backward(self.handle, int(P), dlogp, tape, g_airport, g_route,
                                          g_scales, stream), "ba_backward")

    def generate_noise(self, P, seed, noise, stream):
        _check(load_library().ba_generate_noise(self.handle, int(P), C.c_uint64(int(seed)),
                                                noise, stream), "ba_generate_noise")

    def backward_values(self, P, dlogp, vtape, g_noise, stream):
        _check(load_library().ba_backward_values(self.handle, int(P), dlogp, vtape, g_noise, stream),
               "ba_backward_values")

    def mf_work_bytes(self, P: int) -> int:
        return int(load_library().ba_mf_work_bytes(self.handle, int(P)))

    def mf_sample(self, P, params, seed, d_seed, particle_offset, lat_airport, lat_route, lp0, z,
                  stream):
        _check(load_library().ba_mf_sample(self.handle, int(P), params, C.c_uint64(int(seed)),
                                           d_seed, int(particle_offset), lat_airport, lat_route,
                                           lp0, z, stream), "ba_mf_sample")

    def mf_backward(self, P, params, seed, d_seed, particle_offset, tape, logp, lp0,
                    inv_count_flights, prior_weight, work, flat, stream):
        _check(load_library().ba_mf_backward(self.handle, int(P), params, C.c_uint64(int(seed)),
                                             d_seed, int(particle_offset), tape, logp, lp0,
                                             C.c_float(float(inv_count_flights)),
                                             C.c_float(float(prior_weight)), work, flat,
                                             stream), "ba_mf_backward")

    def logjoint_host(self, P, lat_airport, lat_route, scales, noise, opts, logp, status,
                      g_airport, g_route, g_scales):
        _check(load_library().ba_logjoint_host(self.handle, int(P), lat_airport, lat_route,
                                               scales, noise, C.byref(opts), logp, status,
                                               g_airport, g_route, g_scales),
               "ba_logjoint_host")

    def close(self):
        if getattr(self, "_h", None) is not None and _lib is not None:
            _lib.ba_plan_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _tensor_stream(t):
    import torch
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def rqs_forward(x, params, y, logdet, bins: int, bound: float):
    """``ba_rqs_forward`` on contiguous fp32 CUDA tensors (torch's current stream)."""
    import torch
    with torch.cuda.device(x.device):
        _check(load_library().ba_rqs_forward(x.data_ptr(), params.data_ptr(), y.data_ptr(),
                                             logdet.data_ptr(), x.numel(), int(bins), float(bound),
                                             _tensor_stream(x)), "ba_rqs_forward")


def rqs_backward(x, params, gy, gld, gx, gparams, bins: int, bound: float):
    """``ba_rqs_backward`` on contiguous fp32 CUDA tensors (torch's current stream)."""
    import torch
    with torch.cuda.device(x.device):
        _check(load_library().ba_rqs_backward(x.data_ptr(), params.data_ptr(), gy.data_ptr(),
                                              gld.data_ptr(), gx.data_ptr(), gparams.data_ptr(),
                                              x.numel(), int(bins), float(bound), _tensor_stream(x)),
               "ba_rqs_backward")
