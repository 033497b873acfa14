"""TEST INFRASTRUCTURE ONLY: ctypes front end of the host emulation of the CUDA
forward kernel (tests/hostemu/ba_hostemu.cpp).  Not reachable from the product."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
_OUT = os.path.join(_ROOT, "tests", "_build", "libba_hostemu.so")
_SRCS = [os.path.join(_HERE, "ba_hostemu.cpp"),
         os.path.join(_ROOT, "bayesair_b200", "csrc", "ba_plan_build.cpp")]
_DEPS = _SRCS + [os.path.join(_ROOT, "bayesair_b200", "csrc", f)
                 for f in ("ba_sim.cuh", "ba_scan2.cuh", "ba_plan.h", "ba_plan_build.h")]
_lib = None


class DayTable(C.Structure):
    _fields_ = [("n_airports", C.c_int32), ("n_flights", C.c_int32),
                ("h_sched_dep", C.c_void_p), ("h_act_dep", C.c_void_p),
                ("h_act_arr", C.c_void_p), ("h_cancel_obs", C.c_void_p),
                ("h_origin", C.c_void_p), ("h_dest", C.c_void_p),
                ("h_airport_global", C.c_void_p)]


def build():
    os.makedirs(os.path.dirname(_OUT), exist_ok=True)
    if (not os.path.exists(_OUT)
            or any(os.path.getmtime(s) > os.path.getmtime(_OUT) for s in _DEPS)):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-fPIC",
                               "-shared", "-pthread", "-Wno-unknown-pragmas", "-o", _OUT] + _SRCS)
    return _OUT


def _load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def day_tables(specs):
    keep, arr = [], (DayTable * len(specs))()
    for s, t in zip(specs, arr):
        a = dict(sd=np.ascontiguousarray(s.sched_dep, np.float32),
                 ad=np.ascontiguousarray(s.act_dep, np.float32),
                 aa=np.ascontiguousarray(s.act_arr, np.float32),
                 co=np.ascontiguousarray(s.cancel_obs, np.int8),
                 o=np.ascontiguousarray(s.origin, np.int32),
                 d=np.ascontiguousarray(s.dest, np.int32),
                 g=np.ascontiguousarray(s.airport_global, np.int32))
        keep.append(a)
        t.n_airports, t.n_flights = s.n_airports, s.n_flights
        t.h_sched_dep, t.h_act_dep, t.h_act_arr = a["sd"].ctypes.data, a["ad"].ctypes.data, a["aa"].ctypes.data
        t.h_cancel_obs, t.h_origin, t.h_dest = a["co"].ctypes.data, a["o"].ctypes.data, a["d"].ctypes.data
        t.h_airport_global = a["g"].ctypes.data
    return arr, keep


class EmuPlan:
    def __init__(self, specs, n_airports, delta_t=0.2, max_t=30.0, include_cancellations=False,
                 max_ticks=20000):
        lib = _load()
        arr, self._keep = day_tables(specs)
        self.h = C.c_void_p()
        rc = lib.ba_emu_plan_create(arr, len(specs), n_airports, C.c_float(delta_t),
                                    C.c_float(max_t), int(include_cancellations), max_ticks,
                                    C.byref(self.h))
        if rc != 0:
            raise RuntimeError(f"emu plan create failed: {rc}")
        info = (C.c_int * 6)()
        lib.ba_emu_plan_info(self.h, info)
        self.D, self.N, self.Fmax, self.R, self.C, self.stride = list(info)
        self.route_origin = np.zeros(self.R, np.int32)
        self.route_dest = np.zeros(self.R, np.int32)
        lib.ba_emu_plan_routes(self.h, self.route_origin.ctypes.data_as(C.c_void_p),
                               self.route_dest.ctypes.data_as(C.c_void_p))

    def __del__(self):
        if getattr(self, "h", None):
            _load().ba_emu_plan_destroy(self.h)
            self.h = None

    def pack_latents(self, latents):
        """dict of [P,N] / [P,N,N] arrays -> (lat_airport [P,C,N], lat_route [P,R])."""
        rows = [latents["mean_turnaround"], latents["mean_service"]]
        if self.C == 4:
            rows += [latents["log_initial_aircraft"], latents["base_cancel_logprob"]]
        la = np.ascontiguousarray(np.stack(rows, 1), np.float32)
        lr = np.ascontiguousarray(latents["travel"][:, self.route_origin, self.route_dest], np.float32)
        return la, lr

    def forward(self, latents, scales, noise, *, seed=0, value_mode=False, want_grad=True,
                exact=False, rerun_overflow=True, shuffle_seed=1, want_pop=False,
                replay_waits=False, scan2=False, want_value_grad=False):
        lib = _load()
        la, lr = self.pack_latents(latents)
        P = la.shape[0]
        sc = np.ascontiguousarray(scales, np.float32)
        nz = None if noise is None else np.ascontiguousarray(noise, np.float32)
        if nz is not None:
            assert nz.shape == (P, self.D, self.Fmax, 4), nz.shape
        logp = np.zeros((P, self.D), np.float32)
        status = np.zeros((P, self.D), np.uint32)
        ticks = np.zeros((P, self.D), np.int32)
        tape = np.zeros((P, self.D, self.stride), np.float32) if want_grad else None
        pop = np.zeros((P, self.D, self.Fmax, 2), np.int32) if want_pop else None
        vtape = np.zeros((P, self.D, self.Fmax, 4), np.float32) if want_value_grad else None
        vp = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)  # noqa: E731
        if scan2:
            rc = lib.ba_emu_forward2(self.h, P, vp(la), vp(lr), vp(sc), vp(nz), C.c_uint64(seed),
                                     int(value_mode), int(want_grad), int(rerun_overflow),
                                     C.c_uint(shuffle_seed), vp(logp), vp(status), vp(ticks),
                                     vp(tape), vp(pop), vp(vtape))
            if rc != 0:
                raise RuntimeError("the ring-less scan cannot run this plan")
        else:
            lib.ba_emu_forward(self.h, P, vp(la), vp(lr), vp(sc), vp(nz), C.c_uint64(seed),
                               int(value_mode), int(want_grad), int(exact), int(rerun_overflow),
                               int(replay_waits), C.c_uint(shuffle_seed), vp(logp), vp(status),
                               vp(ticks), vp(tape), vp(pop))
        out = dict(logp=logp, status=status, ticks=ticks, pop_tick=pop, value_grad=vtape)
        if want_grad:
            dl = np.ones((P, self.D), np.float32)
            ga = np.zeros((P, self.C, self.N), np.float32)
            gr = np.zeros((P, self.R), np.float32)
            gsc = np.zeros((P, 3), np.float32)
            lib.ba_emu_backward(self.h, P, vp(dl), vp(tape), vp(ga), vp(gr), vp(gsc))
            gt = np.zeros((P, self.N, self.N), np.float32)
            gt[:, self.route_origin, self.route_dest] = gr
            g = dict(mean_turnaround=ga[:, 0], mean_service=ga[:, 1], travel=gt, scales=gsc)
            if self.C == 4:
                g["log_initial_aircraft"] = ga[:, 2]
                g["base_cancel_logprob"] = ga[:, 3]
            out["grads"] = g
        return out

    @property
    def scan2_ok(self):
        return bool(_load().ba_emu_scan2_ok(self.h))

    def philox_noise(self, P, seed):
        nz = np.zeros((P, self.D, self.Fmax, 4), np.float32)
        _load().ba_emu_noise(self.h, P, C.c_uint64(seed), nz.ctypes.data_as(C.c_void_p))
        return nz


def _predict(self, latents, scales, noise, noise2, *, seed=0, shuffle_seed=1, want_pop=False):
    """Posterior-predictive run: returns dict(sim [P,D,Fmax,3], status, ticks, pop_tick)."""
    lib = _load()
    la, lr = self.pack_latents(latents)
    P = la.shape[0]
    sc = np.ascontiguousarray(scales, np.float32)
    nz = None if noise is None else np.ascontiguousarray(noise, np.float32)
    n2 = None if noise2 is None else np.ascontiguousarray(noise2, np.float32)
    sim = np.zeros((P, self.D, self.Fmax, 3), np.float32)
    status = np.zeros((P, self.D), np.uint32)
    ticks = np.zeros((P, self.D), np.int32)
    pop = np.zeros((P, self.D, self.Fmax, 2), np.int32) if want_pop else None
    vp = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)  # noqa: E731
    lib.ba_emu_predict(self.h, P, vp(la), vp(lr), vp(sc), vp(nz), vp(n2), C.c_uint64(seed),
                       C.c_uint(shuffle_seed), vp(sim), vp(status), vp(ticks), vp(pop))
    return dict(sim=sim, status=status, ticks=ticks, pop_tick=pop)


EmuPlan.predict = _predict
