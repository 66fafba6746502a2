#!/usr/bin/env python
"""Golden vectors for the feature-engineering path.

Runs the reference's OWN feature-engineering statements -- the lines of
`/root/reference/data-processing/merge_state_actions.py` from `df["alert"] = df.issue.notnull()`
(:114) to `df["remaining_budget"] = ...` (:196), unmodified, extracted from the file and exec'd --
on a small synthetic (fips, date) frame, and stores inputs and outputs in
`tests/golden/feature_engineering.npz`.  The script's `main()` as a whole cannot run here (hydra,
patsy, holidays, tqdm and the raw inputs are absent), but these statements only need pandas/numpy
and a `holidays.US(years=...)` mapping, for which a stub built from `us_summer_holidays` below is
injected (the device path takes the holiday flag as a host-supplied input column as well).  The
B-spline block (:199-210) needs patsy and is NOT covered here; see oracle/feature_oracle.py.

Run where /root/reference exists:  python tests/golden/make_feature_golden.py
"""
import datetime as dt
import os
import sys
import textwrap
import types

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
REF = "/root/reference/data-processing/merge_state_actions.py"
FIRST, LAST = 114, 196  # 1-based, inclusive


def us_summer_holidays(years):
    """Federal holidays between May and September: Memorial Day (last Monday of May), Independence Day
    (and its observed weekday), Labor Day (first Monday of September), Juneteenth from 2021."""
    out = set()
    for y in years:
        d = dt.date(y, 5, 31)
        out.add(d - dt.timedelta(days=d.weekday()))
        j4 = dt.date(y, 7, 4)
        out.add(j4)
        if j4.weekday() == 5:
            out.add(j4 - dt.timedelta(days=1))
        if j4.weekday() == 6:
            out.add(j4 + dt.timedelta(days=1))
        d = dt.date(y, 9, 1)
        out.add(d + dt.timedelta(days=(7 - d.weekday()) % 7))
        if y >= 2021:
            out.add(dt.date(y, 6, 19))
    return out


def synthetic_frame(seed=0):
    """(fips, date)-sorted rows: 7 counties, ragged years (a county without 2007, a summer cut short, one that
    starts late), heat index with many ties, alerts in streaks, issued_in_advance NaN where there is no alert."""
    rng = np.random.default_rng(seed)
    rows = []
    fips_all = ["01001", "04013", "06037", "12086", "36061", "48201", "53033"]
    for ci, f in enumerate(fips_all):
        for y in (2006, 2007, 2008, 2009):
            if ci == 1 and y == 2007:
                continue
            days = pd.date_range(f"{y}-05-01", f"{y}-09-30", freq="D")
            if ci == 2 and y == 2008:
                days = days[:120]
            if ci == 3 and y == 2006:
                days = days[10:]
            base = 24 + 8 * np.sin(np.linspace(0, np.pi, len(days))) + rng.normal(0, 3, len(days))
            hi = np.round(base, 1 if ci % 2 else 0)  # coarse rounding -> plenty of ties for the average rank
            alert = np.zeros(len(days), bool)
            k = 0
            while k < len(days):
                if rng.random() < 0.04:
                    run = int(rng.integers(1, 6))
                    alert[k:k + run] = True
                    k += run
                k += 1
            for d, h, a in zip(days, hi, alert):
                rows.append((f, d, float(h), a, float(rng.integers(0, 5)) / 2 if a else np.nan,
                             ("W" if rng.random() < 0.3 else "Y") if a else None))
    df = pd.DataFrame(rows, columns=["fips", "date", "HImax_C", "_alert", "issued_in_advance", "significance"])
    df["issue"] = np.where(df["_alert"], pd.Timestamp("2000-01-01"), pd.NaT)  # :114 only looks at notnull()
    df["issue"] = pd.to_datetime(df["issue"])
    return df.drop(columns="_alert").sort_values(["fips", "date"]).reset_index(drop=True)


def main():
    with open(REF) as fh:
        lines = fh.readlines()
    assert lines[FIRST - 1].strip() == 'df["alert"] = df.issue.notnull()', lines[FIRST - 1]
    assert lines[LAST - 1].strip() == 'df["remaining_budget"] = df["budget"] - df["rolling_alerts"]', lines[LAST - 1]
    code = textwrap.dedent("".join(lines[FIRST - 1:LAST]))
    df = synthetic_frame()
    inputs = {"fips": df.fips.to_numpy(str), "date_days": df.date.to_numpy("datetime64[D]").astype(np.int64).astype(np.int32),
              "himax_c": df.HImax_C.to_numpy(np.float64), "alert": df.issue.notnull().to_numpy(),
              "issued_in_advance": df.issued_in_advance.to_numpy(np.float64)}
    stub = types.ModuleType("holidays")
    stub.US = lambda years: {d: "holiday" for d in us_summer_holidays(years)}
    ns = {"np": np, "pd": pd, "dt": dt, "holidays": stub, "df": df}
    exec(compile(code, REF, "exec"), ns)  # the reference's statements, as they stand
    df = ns["df"]
    cols = ["hi_max", "heat_qi", "heat_qi_above_25", "heat_qi_above_75", "hi_max_above_25", "hi_max_above_75",
            "hi_max*heat_qi", "hi_max_above_25*heat_qi", "hi_max_above_75*heat_qi", "heat_qi_3d", "heat_qi_7d",
            "excess_heat_3d", "excess_heat_7d", "excess_heat_3d*heat_qi", "excess_heat_7d*heat_qi", "year", "dos",
            "alerts_2wks", "alert_lag1", "alert_streak", "issued_in_advance", "weekend", "holiday", "budget",
            "rolling_alerts", "remaining_budget"]
    out = {f"in_{k}": v for k, v in inputs.items()}
    out["in_holiday"] = df["holiday"].to_numpy(np.uint8)  # host-supplied input of the device path
    for c in cols:
        out["out_" + c] = df[c].to_numpy()
    out["out_date_str"] = df["date"].to_numpy(str)
    dst = os.path.join(HERE, "feature_engineering.npz")
    np.savez_compressed(dst, **out)
    print(f"{len(df)} rows, {df.fips.nunique()} counties -> {dst}", file=sys.stderr)
    print(df[cols].describe().T[["min", "max"]].to_string(), file=sys.stderr)


if __name__ == "__main__":
    main()
