"""bayesair_b200 -- the hot path of dawsonc/BayesAir on B200 (sm_100a).

Public surface (mirrors ``bayes_air``):
    air_traffic_network_model, NetworkState, Flight, Airport, QueueEntry,
    parse_flight, parse_schedule
plus the batched entry points the kernels make possible:
    SimulationPlan, simulate_log_joint, flight_noise
"""
from .network import NetworkState  # noqa: F401
from .schedule import parse_flight, parse_schedule  # noqa: F401
from .types import Airport, AirportCode, Flight, QueueEntry, Time  # noqa: F401

__version__ = "0.1.0"


def __getattr__(name):
    # the CUDA-backed pieces import torch + ctypes lazily so that the pure-host
    # types above stay importable in tools that only build schedules
    if name in ("air_traffic_network_model", "flight_noise", "FACTOR_NAME"):
        from . import model
        return getattr(model, name)
    if name in ("SimulationPlan", "plan_for_states", "clear_plan_cache"):
        from . import plan
        return getattr(plan, name)
    if name == "simulate_log_joint":
        from .function import simulate_log_joint
        return simulate_log_joint
    raise AttributeError(name)
