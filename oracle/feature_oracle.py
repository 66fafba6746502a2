"""CPU restatement of the reference's feature engineering (TEST INFRASTRUCTURE, not the product).

Follows `/root/reference/data-processing/merge_state_actions.py:114-210` -- the statements that
turn (fips, date, HImax_C, alert, issued_in_advance) rows into the feature-store columns the env
reads -- in plain numpy, one county at a time, every block citing the lines it restates.

Pinned: `tests/golden/feature_engineering.npz` holds the outputs of the reference's OWN statements
(:114-196, extracted from the file and exec'd by `tests/golden/make_feature_golden.py`);
`tests/test_feature_oracle.py` replays them.  Integer columns agree exactly, float64 columns to
1e-12 (pandas' rolling mean is an online Kahan sum, `pandas/_libs/window/aggregations.pyx`, whose
last bits differ from a direct window sum).

NOT pinned against the reference: the B-spline block (:199-210) calls `patsy.dmatrix`, and patsy
(like `holidays`) is not installed in this image and not vendored by the reference
(`pyproject.toml:20-29` lists neither; `requirements.yaml` pulls them from conda).  `bspline_dos`
restates patsy's published algorithm (`patsy/splines.py`, `bs`: df=3, degree=3, no intercept ->
no interior knots, boundary knots repeated 4x, first basis column dropped) and is cross-checked
against `scipy.interpolate.BSpline` -- "parity unpinned" for those three columns.
"""
from __future__ import annotations

import numpy as np

FLOAT_COLUMNS = ("hi_max", "heat_qi", "heat_qi_above_25", "heat_qi_above_75", "hi_max_above_25", "hi_max_above_75",
                 "hi_max*heat_qi", "hi_max_above_25*heat_qi", "hi_max_above_75*heat_qi", "heat_qi_3d", "heat_qi_7d",
                 "excess_heat_3d", "excess_heat_7d", "excess_heat_3d*heat_qi", "excess_heat_7d*heat_qi",
                 "alerts_2wks", "issued_in_advance")
INT_COLUMNS = ("year", "dos", "alert_lag1", "alert_streak", "weekend", "budget", "rolling_alerts", "remaining_budget")


def civil_year(days: np.ndarray) -> np.ndarray:
    """Calendar year of days-since-1970-01-01 (`df.date.dt.year`, :150)."""
    return (np.asarray(days).astype("datetime64[D]").astype("datetime64[Y]").astype(np.int64) + 1970).astype(np.int32)


def rank_pct(v: np.ndarray) -> np.ndarray:
    """`Series.rank(pct=True)` (:122): average rank of ties over the number of valid values, NaN kept."""
    ok = ~np.isnan(v)
    out = np.full(v.shape, np.nan)
    x = v[ok]
    less = (x[None, :] < x[:, None]).sum(axis=1)
    equal = (x[None, :] == x[:, None]).sum(axis=1)
    out[ok] = (less + (equal + 1) * 0.5) / max(len(x), 1)
    return out


def rolling_mean(v: np.ndarray, w: int) -> np.ndarray:
    """`x.rolling(w, min_periods=1).mean()` (:136-141)."""
    out = np.empty_like(v)
    for i in range(len(v)):
        win = v[max(0, i - w + 1):i + 1]
        win = win[~np.isnan(win)]
        out[i] = win.sum() / len(win) if len(win) else np.nan
    return out


def engineer(county_start, date_days, himax_c, alert, issued_in_advance):
    """Rows sorted by (county, date); `county_start[c] : county_start[c+1]` are county c's rows.
    Returns a dict column -> array over all rows (float64 / int32)."""
    n = len(date_days)
    out = {k: np.zeros(n, np.float64) for k in FLOAT_COLUMNS}
    out.update({k: np.zeros(n, np.int32) for k in INT_COLUMNS})
    himax_c = np.asarray(himax_c, np.float64)
    alert = np.asarray(alert).astype(bool)
    out["hi_max"][:] = 0.01 * (himax_c * 9 / 5 + 32)                         # :121
    out["year"][:] = civil_year(date_days)                                    # :150
    out["weekend"][:] = ((np.asarray(date_days, np.int64) + 3) % 7) >= 5      # :178 (1970-01-01 is a Thursday)
    out["issued_in_advance"][:] = np.nan_to_num(np.asarray(issued_in_advance, np.float64), nan=0.0)  # :175
    for c in range(len(county_start) - 1):
        sl = slice(int(county_start[c]), int(county_start[c + 1]))
        hi, a, yr = out["hi_max"][sl], alert[sl], out["year"][sl]
        q = rank_pct(hi)                                                      # :122 groupby("fips").rank(pct=True)
        out["heat_qi"][sl] = q
        q3, q7 = rolling_mean(q, 3), rolling_mean(q, 7)                       # :136-141 (rolls across years)
        out["heat_qi_3d"][sl], out["heat_qi_7d"][sl] = q3, q7
        m = len(a)
        out["alerts_2wks"][sl] = [a[max(0, i - 13):i + 1].sum() for i in range(m)]   # :154-156
        out["alert_lag1"][sl] = np.r_[0, a[:-1]].astype(np.int32)             # :157
        streak, s = np.zeros(m, np.int32), 0
        for i in range(m):                                                     # :161-172
            s = s + 1 if a[i] else 0
            streak[i] = s
        out["alert_streak"][sl] = streak
        for y in np.unique(yr):                                                # groups of (fips, year)
            g = np.flatnonzero(yr == y)
            out["dos"][sl][g] = np.arange(len(g))                             # :151 rank of the date inside the group - 1
            out["budget"][sl][g] = a[g].sum()                                  # :190-192
            out["rolling_alerts"][sl][g] = np.cumsum(a[g])                     # :195
    q, hi = out["heat_qi"], out["hi_max"]
    out["heat_qi_above_25"][:] = (q > 0.25).astype(np.int64) * q             # :125-128
    out["heat_qi_above_75"][:] = (q > 0.75).astype(np.int64) * q
    out["hi_max_above_25"][:] = (hi > 25).astype(np.int64) * hi
    out["hi_max_above_75"][:] = (hi > 75).astype(np.int64) * hi
    out["hi_max*heat_qi"][:] = q * hi                                         # :131-133
    out["hi_max_above_25*heat_qi"][:] = out["heat_qi_above_25"] * hi
    out["hi_max_above_75*heat_qi"][:] = out["heat_qi_above_75"] * hi
    out["excess_heat_3d"][:] = np.clip(q - out["heat_qi_3d"], 0, None)        # :144-147
    out["excess_heat_7d"][:] = np.clip(q - out["heat_qi_7d"], 0, None)
    out["excess_heat_3d*heat_qi"][:] = out["excess_heat_3d"] * q
    out["excess_heat_7d*heat_qi"][:] = out["excess_heat_7d"] * q
    out["remaining_budget"][:] = out["budget"] - out["rolling_alerts"]        # :196
    return out


def bernstein3(u: np.ndarray) -> np.ndarray:
    """The cubic B-spline basis on knots [0,0,0,0,1,1,1,1] without its first function: columns
    3u(1-u)^2, 3u^2(1-u), u^3 -- what `bs(x, df=3, degree=3)` evaluates with no interior knots."""
    u = np.asarray(u, np.float64)
    return np.stack([3 * u * (1 - u) ** 2, 3 * u ** 2 * (1 - u), u ** 3], axis=1)


def bspline_dos(dos: np.ndarray):
    """:198-210: `bs(dos / M, df=3, degree=3, lower_bound=0, upper_bound=M + 1) - 1`, every column standardised with
    its mean and (ddof=1) standard deviation over all rows.  Returns (columns [n,3], means[3], stds[3])."""
    dos = np.asarray(dos, np.float64)
    M = dos.max()
    basis = bernstein3((dos / M) / (M + 1))  # the bounds are [0, M+1] although dos/M lies in [0, 1] (as in the reference)
    mean = basis.mean(axis=0)
    std = basis.std(axis=0, ddof=1)
    return (basis - mean) / std, mean, std
