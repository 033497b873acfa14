"""Schedule parsing: dataframe rows -> ``Flight`` / ``Airport`` objects.

Same entry points and column contract as the reference's
``bayes_air/schedule.py:9-106`` (``parse_flight``, ``parse_schedule``): times are
float hours and become 0-d fp32 tensors, a cancelled row loses its four
measured times (they become ``None``), ``actually_cancelled`` is a 0/1 fp32
tensor, and the airport list is in order of first appearance in
``concat([origin column, destination column])``.
"""
from __future__ import annotations

from typing import List, Tuple

import pandas as pd
import torch

from .types import Airport, Flight

_MEASURED = ("actual_departure_time", "actual_arrival_time",
             "wheels_off_time", "wheels_on_time")


def _as_time(x, device):
    return None if x is None else torch.tensor(float(x), dtype=torch.float32, device=device)


def parse_flight(schedule_row, device=None) -> Flight:
    device = torch.device("cpu") if device is None else device
    is_cancelled = bool(schedule_row["cancelled"])
    measured = {k: (None if is_cancelled else schedule_row[k]) for k in _MEASURED}
    return Flight(
        flight_number=schedule_row["flight_number"],
        origin=schedule_row["origin_airport"],
        destination=schedule_row["destination_airport"],
        scheduled_departure_time=_as_time(schedule_row["scheduled_departure_time"], device),
        scheduled_arrival_time=_as_time(schedule_row["scheduled_arrival_time"], device),
        actually_cancelled=torch.tensor(1.0 if is_cancelled else 0.0, device=device),
        actual_departure_time=_as_time(measured["actual_departure_time"], device),
        actual_arrival_time=_as_time(measured["actual_arrival_time"], device),
        wheels_off_time=_as_time(measured["wheels_off_time"], device),
        wheels_on_time=_as_time(measured["wheels_on_time"], device),
    )


def parse_schedule(schedule_df: pd.DataFrame, device=None) -> Tuple[List[Flight], List[Airport]]:
    records = schedule_df.to_dict("records")
    flights = [parse_flight(row, device=device) for row in records]
    codes = pd.unique(
        pd.concat([schedule_df["origin_airport"], schedule_df["destination_airport"]])
    )
    return flights, [Airport(str(c)) for c in codes]
