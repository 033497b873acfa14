"""Boundary types of the hot path: ``Flight``, ``Airport``, ``QueueEntry``.

These mirror the *input contract* of the reference model
(``bayes_air/types/flight.py:10-45``, ``bayes_air/types/airport.py:13-64``,
``bayes_air/types/util.py:5-6``): same class names, same field names, same
``str(flight)`` (which is the stem of every per-flight sample-site name), so a
caller that builds ``NetworkState`` objects for the reference can hand them to
``bayesair_b200.air_traffic_network_model`` unchanged.

Unlike the reference, these classes carry *no simulation methods*: the queue
update / service / turnaround logic (``airport.py:66-221``) lives in the CUDA
kernel (``csrc/ba_sim.cuh``).  Objects of the reference's own classes are
accepted too -- the schedule compiler only reads attributes (duck typing).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import torch

AirportCode = str
Time = torch.tensor  # the reference's alias (types/util.py:6): a 0-d fp32 tensor


def _t(x: float):
    return lambda: Time(x)


@dataclass
class Flight:
    """One scheduled leg.  Field-for-field the reference's ``Flight``."""

    flight_number: str
    origin: AirportCode
    destination: AirportCode
    scheduled_departure_time: torch.Tensor
    scheduled_arrival_time: torch.Tensor
    actually_cancelled: Optional[torch.Tensor]
    simulated_cancelled: Optional[torch.Tensor] = None
    simulated_departure_time: Optional[torch.Tensor] = None
    simulated_arrival_time: Optional[torch.Tensor] = None
    actual_departure_time: Optional[torch.Tensor] = None
    actual_arrival_time: Optional[torch.Tensor] = None
    wheels_on_time: Optional[torch.Tensor] = None
    wheels_off_time: Optional[torch.Tensor] = None

    def __str__(self) -> str:  # site-name stem, flight.py:44-45
        return "_".join((str(self.flight_number), self.origin, self.destination))


@dataclass
class QueueEntry:
    """Kept for API compatibility (airport.py:13-27); unused by the CUDA path."""

    flight: Flight
    queue_start_time: torch.Tensor
    total_wait_time: torch.Tensor = field(default_factory=_t(0.0))
    assigned_service_time: Optional[torch.Tensor] = None


@dataclass
class Airport:
    """One airport of a network day.  Only ``code`` is read by the compiler; the
    remaining fields exist so that code written against the reference's
    ``Airport`` (airport.py:30-64) keeps working."""

    code: AirportCode
    mean_service_time: torch.Tensor = field(default_factory=_t(1.0))
    runway_use_time_std_dev: torch.Tensor = field(default_factory=_t(1.0))
    mean_turnaround_time: torch.Tensor = field(default_factory=_t(1.0))
    turnaround_time_std_dev: torch.Tensor = field(default_factory=_t(1.0))
    num_available_aircraft: torch.Tensor = field(default_factory=_t(0.0))
    base_cancel_prob: torch.Tensor = field(default_factory=_t(0.0))
    runway_queue: List[QueueEntry] = field(default_factory=list)
    turnaround_queue: List[torch.Tensor] = field(default_factory=list)
    available_aircraft: List[torch.Tensor] = field(default_factory=list)
    last_departure_time: torch.Tensor = field(default_factory=_t(0.0))


__all__ = ["AirportCode", "Time", "Flight", "Airport", "QueueEntry"]
