"""On-disk flight data -> the dataframe schema the schedule compiler consumes.

Host-side mirror of the reference's ``bayes_air/utils/dataloader.py`` (SURVEY.md §8 row
f4): same function names, argument meaning and results, so that

    df, locations = load_all_data(data_dir)
    df = remap_columns(df, locations)
    days = split_by_date(top_N_df(df, 100))
    specs = [compiler.day_spec_from_frame(d) for d in days]

feeds ``SimulationPlan`` exactly what the reference's ``parse_schedule`` would see.
What is kept on purpose, because the model's inputs depend on it:

* times are wall-clock ``HH:MM`` strings localised on the date pandas gives a bare time
  (1900-01-01) and converted to America/Denver, hour + minute/60, date dropped
  (``dataloader.py:88-120``).  The conversion therefore uses each zone's 1900 offset (no
  daylight saving; zones that were on local mean time then, e.g. America/Detroit, keep that
  quirk) -- it is done with the same pandas calls as the reference, once per time zone
  instead of once per row;
* ``"--:--"`` (no time recorded) and a genuine ``00:00`` both become ``max + 1``
  (``dataloader.py:100-101,116-117``);
* ``wheels_on_time`` is converted with the *origin* airport's zone (``dataloader.py:224-227``);
* the +24 h repairs for flights that cross midnight or leave a day late
  (``dataloader.py:229-271``), including their order.

``time_zone_for_airports`` uses ``timezonefinder`` when it is installed (the reference's
dependency); otherwise a table from the airport's state (with the usual split-state
exceptions), which is exact for every airport Southwest serves.
"""
from __future__ import annotations

import os
from typing import Iterable, List, Optional, Tuple

import pandas as pd

__all__ = [
    "load_all_data", "split_nominal_disrupted_data", "split_by_date",
    "convert_to_float_hours_optimized", "time_zone_for_airports", "remap_columns",
    "top_N_df", "COLUMN_MAPPING",
]

# reference: dataloader.py:159-171
COLUMN_MAPPING = {
    "Flight Number": "flight_number",
    "Date": "date",
    "Origin Airport Code": "origin_airport",
    "Dest Airport Code": "destination_airport",
    "Scheduled Departure Time": "scheduled_departure_time",
    "Scheduled Arrival Time": "scheduled_arrival_time",
    "Actual Departure Time": "actual_departure_time",
    "Actual Arrival Time": "actual_arrival_time",
    "Wheels On Time": "wheels_on_time",
    "Wheels Off Time": "wheels_off_time",
    "Cancelled Flight": "cancelled",
}

_REFERENCE_ZONE = "America/Denver"


def _default_data_dir() -> str:
    return os.environ.get("BAYESAIR_DATA_DIR", os.path.join(os.getcwd(), "data"))


def load_all_data(data_dir: Optional[str] = None) -> Tuple[pd.DataFrame, pd.DataFrame]:
    """All available WN flight data + the latest airport locations
    (``dataloader.py:8-40``).  ``data_dir`` (default ``$BAYESAIR_DATA_DIR`` or ``./data``)
    holds ``wn_dec01_dec20.csv``, ``wn_dec21_dec30.csv`` and ``airport_locations.csv``."""
    data_dir = data_dir or _default_data_dir()
    nominal = pd.read_csv(os.path.join(data_dir, "wn_dec01_dec20.csv"))
    disrupted = pd.read_csv(os.path.join(data_dir, "wn_dec21_dec30.csv"))
    locations = pd.read_csv(os.path.join(data_dir, "airport_locations.csv"))
    locations = locations[locations.AIRPORT_IS_LATEST.astype(bool)]
    df = pd.concat([nominal, disrupted]).drop_duplicates()
    return df, locations


def split_nominal_disrupted_data(df: pd.DataFrame):
    """(flights outside, flights inside) the 2022-12-21 ... 2022-12-30 disruption
    (``dataloader.py:43-63``)."""
    start, end = pd.to_datetime("12/21/2022"), pd.to_datetime("12/30/2022")
    nominal = df[(df["date"] < start) | (df["date"] > end)]
    disrupted = df[(df["date"] >= start) & (df["date"] <= end)]
    return nominal, disrupted


def split_by_date(df: pd.DataFrame) -> List[pd.DataFrame]:
    """One dataframe per date, each sorted by scheduled departure
    (``dataloader.py:66-85``; the sort is pandas' default, as there)."""
    frames = [group for _, group in df.groupby("date")]
    return [g.sort_values(by="scheduled_departure_time") for g in frames]


def convert_to_float_hours_optimized(time_series: pd.Series, time_zone_series: pd.Series) -> pd.Series:
    """``HH:MM`` wall-clock strings in the given zones -> float hours since midnight in
    America/Denver (``dataloader.py:88-120``).  Missing (``--:--``) and ``00:00`` entries
    become the largest converted value + 1."""
    ts = time_series.replace("--:--", "00:00").replace("24:00", "23:59")
    stamps = pd.to_datetime(ts, format="%H:%M")
    hours = pd.Series(0.0, index=ts.index, dtype="float64")
    zones = time_zone_series.reindex(ts.index) if not time_zone_series.index.equals(ts.index) \
        else time_zone_series
    for zone, idx in zones.groupby(zones, sort=False).groups.items():
        local = stamps.loc[idx].dt.tz_localize(zone).dt.tz_convert(_REFERENCE_ZONE)
        hours.loc[idx] = (local.dt.hour + local.dt.minute / 60.0).to_numpy()
    hours[ts == "00:00"] = hours.max() + 1.0
    return hours


# ---- airport -> IANA zone ----------------------------------------------------------
_STATE_ZONE = {
    "AL": "America/Chicago", "AK": "America/Anchorage", "AZ": "America/Phoenix",
    "AR": "America/Chicago", "CA": "America/Los_Angeles", "CO": "America/Denver",
    "CT": "America/New_York", "DE": "America/New_York", "DC": "America/New_York",
    "FL": "America/New_York", "GA": "America/New_York", "HI": "Pacific/Honolulu",
    "ID": "America/Boise", "IL": "America/Chicago", "IN": "America/Indiana/Indianapolis",
    "IA": "America/Chicago", "KS": "America/Chicago", "KY": "America/New_York",
    "LA": "America/Chicago", "ME": "America/New_York", "MD": "America/New_York",
    "MA": "America/New_York", "MI": "America/Detroit", "MN": "America/Chicago",
    "MS": "America/Chicago", "MO": "America/Chicago", "MT": "America/Denver",
    "NE": "America/Chicago", "NV": "America/Los_Angeles", "NH": "America/New_York",
    "NJ": "America/New_York", "NM": "America/Denver", "NY": "America/New_York",
    "NC": "America/New_York", "ND": "America/Chicago", "OH": "America/New_York",
    "OK": "America/Chicago", "OR": "America/Los_Angeles", "PA": "America/New_York",
    "RI": "America/New_York", "SC": "America/New_York", "SD": "America/Chicago",
    "TN": "America/Chicago", "TX": "America/Chicago", "UT": "America/Denver",
    "VT": "America/New_York", "VA": "America/New_York", "WA": "America/Los_Angeles",
    "WV": "America/New_York", "WI": "America/Chicago", "WY": "America/Denver",
    "PR": "America/Puerto_Rico", "VI": "America/St_Thomas", "TT": "Pacific/Guam",
}
# airports on the other side of a split state (commercial-service airports only)
_AIRPORT_ZONE = {
    "ELP": "America/Denver",                                     # far-west Texas
    "PNS": "America/Chicago", "VPS": "America/Chicago", "ECP": "America/Chicago",   # FL panhandle
    "SDF": "America/Kentucky/Louisville", "PAH": "America/Chicago", "OWB": "America/Chicago",
    "TYS": "America/New_York", "CHA": "America/New_York", "TRI": "America/New_York",
    "EVV": "America/Chicago", "GYY": "America/Chicago",
    "GEG": "America/Los_Angeles", "LWS": "America/Los_Angeles", "COE": "America/Los_Angeles",
    "RAP": "America/Denver", "BIS": "America/Chicago", "DIK": "America/Denver",
    "LBF": "America/Chicago", "BFF": "America/Denver", "GCK": "America/Chicago",
    "IMT": "America/Menominee", "IWD": "America/Menominee", "CMX": "America/Detroit",
}


def _zone_from_table(code: str, airport_locations_df: pd.DataFrame) -> str:
    if code in _AIRPORT_ZONE:
        return _AIRPORT_ZONE[code]
    row = airport_locations_df[airport_locations_df.AIRPORT == code]
    if len(row) == 0:
        raise KeyError(f"airport {code!r} is not in the airport location table")
    state = row.AIRPORT_STATE_CODE.iat[0] if "AIRPORT_STATE_CODE" in row else None
    if isinstance(state, str) and state in _STATE_ZONE:
        return _STATE_ZONE[state]
    raise KeyError(f"no time zone known for airport {code!r} (state {state!r}); "
                   "install timezonefinder")


def time_zone_for_airports(airport_codes: Iterable[str], airport_locations_df: pd.DataFrame) -> List[str]:
    """IANA zone of each airport (``dataloader.py:123-145``)."""
    try:
        import timezonefinder as tzf
    except ImportError:
        tzf = None
    if tzf is not None:
        finder = tzf.TimezoneFinder()
        zones = []
        for code in airport_codes:
            row = airport_locations_df[airport_locations_df.AIRPORT == code]
            zones.append(finder.timezone_at(lng=row.LONGITUDE.iat[0], lat=row.LATITUDE.iat[0]))
        return zones
    return [_zone_from_table(str(code), airport_locations_df) for code in airport_codes]


def remap_columns(df: pd.DataFrame, airport_locations_df: pd.DataFrame) -> pd.DataFrame:
    """BTS column names -> the model's, times -> float hours (``dataloader.py:148-283``)."""
    out = df[list(COLUMN_MAPPING.keys())].rename(columns=COLUMN_MAPPING)
    codes = pd.concat([out["origin_airport"], out["destination_airport"]]).unique()
    zones = pd.DataFrame({"airport_code": codes,
                          "time_zone": time_zone_for_airports(codes, airport_locations_df)})
    out = out.merge(zones, left_on="origin_airport", right_on="airport_code")
    out = out.rename(columns={"time_zone": "origin_time_zone"})
    out = out.merge(zones, left_on="destination_airport", right_on="airport_code")
    out = out.rename(columns={"time_zone": "destination_time_zone"})
    out["cancelled"] = out["cancelled"] == "Yes"

    for col, zone_col in (
        ("scheduled_departure_time", "origin_time_zone"),
        ("actual_departure_time", "origin_time_zone"),
        ("wheels_off_time", "origin_time_zone"),
        ("scheduled_arrival_time", "destination_time_zone"),
        ("actual_arrival_time", "destination_time_zone"),
        ("wheels_on_time", "origin_time_zone"),        # sic: dataloader.py:224-227
    ):
        out[col] = convert_to_float_hours_optimized(out[col], out[zone_col])

    # en route at midnight: durations would be negative (dataloader.py:229-247)
    scheduled_duration = out.scheduled_arrival_time - out.scheduled_departure_time
    actual_duration = out.actual_arrival_time - out.actual_departure_time
    wheels_up_duration = out.wheels_on_time - out.wheels_off_time
    out.loc[actual_duration < 0, "actual_arrival_time"] += 24
    out.loc[wheels_up_duration < 0, "wheels_on_time"] += 24
    out.loc[scheduled_duration < 0, "scheduled_arrival_time"] += 24

    # departed the next day (dataloader.py:249-271): very negative departure delay
    departure_delay = out.actual_departure_time - out.scheduled_departure_time
    late = departure_delay < -3.0
    for col in ("actual_departure_time", "wheels_off_time", "wheels_on_time", "actual_arrival_time"):
        out.loc[late, col] += 24

    out["date"] = pd.to_datetime(out["date"])
    return out


def top_N_df(df: pd.DataFrame, number_of_airports: int) -> pd.DataFrame:
    """Flights between the N airports with the most arrivals (``dataloader.py:286-306``)."""
    top = df["destination_airport"].value_counts().head(number_of_airports).index
    return df[df["origin_airport"].isin(top) & df["destination_airport"].isin(top)]
