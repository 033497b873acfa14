"""Generates tests/golden/dataloader/: a small synthetic data directory in the on-disk
format of the reference (two BTS-style CSVs + an airport location table) and what the
UNMODIFIED reference dataloader (`/root/reference/bayes_air/utils/dataloader.py`) makes of
it.  Build container only (imports the reference); the outputs are committed.

    python tests/golden/make_dataloader_golden.py

`timezonefinder` is absent from this image; the reference imports it at module level, so a
stand-in module is injected whose `timezone_at(lng, lat)` answers from the coordinates of
the synthetic location table (the zones are those of the real airports).
"""
import json
import os
import sys
import types

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "dataloader")
REFERENCE_ROOT = os.environ.get("BAYESAIR_REFERENCE_ROOT", "/root/reference")

# code, state, lat, lon, zone   (public facts about the real airports, rounded)
AIRPORTS = [
    ("DEN", "CO", 39.86, -104.67, "America/Denver"),
    ("DAL", "TX", 32.85, -96.85, "America/Chicago"),
    ("MDW", "IL", 41.79, -87.75, "America/Chicago"),
    ("PHX", "AZ", 33.43, -112.01, "America/Phoenix"),
    ("HOU", "TX", 29.65, -95.28, "America/Chicago"),
    ("LAS", "NV", 36.08, -115.15, "America/Los_Angeles"),
    ("MCO", "FL", 28.43, -81.31, "America/New_York"),
    ("BNA", "TN", 36.12, -86.68, "America/Chicago"),
    ("BWI", "MD", 39.18, -76.67, "America/New_York"),
    ("OAK", "CA", 37.72, -122.22, "America/Los_Angeles"),
    ("ELP", "TX", 31.81, -106.38, "America/Denver"),
    ("DTW", "MI", 42.21, -83.35, "America/Detroit"),
    ("BOI", "ID", 43.56, -116.22, "America/Boise"),
    ("HNL", "HI", 21.32, -157.92, "Pacific/Honolulu"),
]


def hhmm(hours: float) -> str:
    m = int(round(hours * 60)) % (24 * 60)
    return f"{m // 60:02d}:{m % 60:02d}"


def make_raw(dates, rng, n_per_day):
    rows = []
    codes = [a[0] for a in AIRPORTS]
    weights = np.array([8, 7, 7, 6, 6, 6, 5, 5, 5, 4, 1, 1, 1, 1], float)
    weights /= weights.sum()
    fn = 100
    for date in dates:
        for _ in range(n_per_day):
            o, d = rng.choice(len(codes), size=2, replace=False, p=weights)
            sched = rng.uniform(5.0, 23.5)
            dur = rng.uniform(1.0, 5.5)
            delay = rng.exponential(0.3) - 0.05
            if rng.random() < 0.04:
                delay += rng.uniform(3.0, 9.0)         # some leave the next day
            cancelled = rng.random() < 0.08
            fn += 1
            rows.append({
                "Carrier Code": "WN", "Date": date, "Flight Number": fn,
                "Tail Number": f"N{rng.integers(100, 999)}WN",
                "Origin Airport Code": codes[o], "Dest Airport Code": codes[d],
                "Scheduled Departure Time": hhmm(sched),
                "Scheduled Arrival Time": hhmm(sched + dur),
                "Actual Departure Time": "--:--" if cancelled else hhmm(sched + delay),
                "Wheels Off Time": "--:--" if cancelled else hhmm(sched + delay + 0.2),
                "Wheels On Time": "--:--" if cancelled else hhmm(sched + delay + dur - 0.1),
                "Actual Arrival Time": "--:--" if cancelled else hhmm(sched + delay + dur + rng.normal(0, 0.1)),
                "Cancelled Flight": "Yes" if cancelled else "No",
            })
    # hand-made corner cases: midnight strings, a genuine 00:00
    rows.append({**rows[0], "Flight Number": 9001, "Scheduled Arrival Time": "24:00",
                 "Actual Arrival Time": "00:00", "Cancelled Flight": "No",
                 "Actual Departure Time": rows[0]["Scheduled Departure Time"],
                 "Wheels Off Time": rows[0]["Scheduled Departure Time"], "Wheels On Time": "23:50"})
    return pd.DataFrame(rows)


def main():
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(2022)
    nominal = make_raw(["12/18/2022", "12/19/2022", "12/20/2022"], rng, 70)
    disrupted = make_raw(["12/21/2022", "12/22/2022"], rng, 70)
    # a duplicated row across the two files (the reference de-duplicates)
    disrupted = pd.concat([disrupted, nominal.iloc[[5]]])
    loc = pd.DataFrame([{
        "AIRPORT_SEQ_ID": 1000000 + i, "AIRPORT_ID": 10000 + i, "AIRPORT": c,
        "AIRPORT_STATE_CODE": st, "LATITUDE": lat, "LONGITUDE": lon,
        "AIRPORT_IS_CLOSED": 0, "AIRPORT_IS_LATEST": 1,
    } for i, (c, st, lat, lon, _) in enumerate(AIRPORTS)])
    # a superseded record that must be filtered out (wrong coordinates on purpose)
    old = loc.iloc[[0]].copy()
    old["AIRPORT_IS_LATEST"] = 0
    old["LATITUDE"], old["LONGITUDE"] = 0.0, 0.0
    loc = pd.concat([old, loc])
    nominal.to_csv(os.path.join(OUT, "wn_dec01_dec20.csv"), index=False)
    disrupted.to_csv(os.path.join(OUT, "wn_dec21_dec30.csv"), index=False)
    loc.to_csv(os.path.join(OUT, "airport_locations.csv"), index=False)

    # ---- the reference, with a stand-in for the absent timezonefinder ----------------
    zone_at = {(round(lat, 2), round(lon, 2)): z for _, _, lat, lon, z in AIRPORTS}

    class TimezoneFinder:
        def timezone_at(self, lng, lat):
            return zone_at[(round(float(lat), 2), round(float(lng), 2))]

    sys.modules["timezonefinder"] = types.SimpleNamespace(TimezoneFinder=TimezoneFinder)
    sys.path.append(REFERENCE_ROOT)
    import importlib
    ref = importlib.import_module("bayes_air.utils.dataloader")

    n_df = pd.read_csv(os.path.join(OUT, "wn_dec01_dec20.csv"))
    d_df = pd.read_csv(os.path.join(OUT, "wn_dec21_dec30.csv"))
    l_df = pd.read_csv(os.path.join(OUT, "airport_locations.csv"))
    l_df = l_df[l_df.AIRPORT_IS_LATEST.astype(bool)]          # load_all_data's filter
    df = pd.concat([n_df, d_df]).drop_duplicates()            # load_all_data's concat
    remapped = ref.remap_columns(df, l_df)
    remapped.to_csv(os.path.join(OUT, "expected_remap.csv"), index=False)
    top = ref.top_N_df(remapped, 6)
    nom, dis = ref.split_nominal_disrupted_data(top)
    days = ref.split_by_date(nom) + ref.split_by_date(dis)
    meta = {
        "pandas": pd.__version__,
        "n_raw": int(len(df)), "n_remapped": int(len(remapped)),
        "top6_airports": sorted(set(top.origin_airport) | set(top.destination_airport)),
        "n_top6": int(len(top)), "n_nominal": int(len(nom)), "n_disrupted": int(len(dis)),
        "days": [{"date": str(d["date"].iloc[0].date()), "n": int(len(d)),
                  "flight_numbers": [int(x) for x in d["flight_number"]]} for d in days],
    }
    json.dump(meta, open(os.path.join(OUT, "expected_meta.json"), "w"), indent=1)
    print(json.dumps({k: v for k, v in meta.items() if k != "days"}, indent=1))
    print(remapped.head())


if __name__ == "__main__":
    main()
