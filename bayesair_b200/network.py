"""``NetworkState``: the per-day input container of the model.

Mirror of the reference's ``bayes_air.network.NetworkState``
(``bayes_air/network.py:12-41``): same constructor, same stable sort of
``pending_flights`` by scheduled departure time in ``__post_init__`` (ties
keep their input order), same ``complete`` property.  The mutating methods of
the reference class (``pop_ready_to_depart_flights``, ``add_in_transit_flights``,
``update_in_transit_flights``; ``network.py:43-198``) are *not* here: that work
is the CUDA kernel's.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Tuple

from .types import Airport, AirportCode, Flight, Time


@dataclass
class NetworkState:
    airports: Dict[AirportCode, Airport]
    pending_flights: List[Flight]
    in_transit_flights: List[Tuple[Flight, Time]] = field(default_factory=list)
    completed_flights: List[Flight] = field(default_factory=list)

    def __post_init__(self):
        # list.sort is stable; key is the fp32 scheduled departure (network.py:30-32)
        self.pending_flights.sort(key=lambda f: float(f.scheduled_departure_time))

    @property
    def complete(self) -> bool:
        if self.pending_flights or self.in_transit_flights:
            return False
        return not any(a.runway_queue for a in self.airports.values())
