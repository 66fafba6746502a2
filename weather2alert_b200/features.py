"""Feature engineering on the device: raw daily heat-index / alert records -> the feature store.

The reference builds its feature store offline with pandas
(`/root/reference/data-processing/merge_state_actions.py:114-210`, run under Snakemake); here the same
columns come out of three kernels behind `w2a_engineer_features` (`csrc/w2a_features.cuh`), so new weather
can be turned into an HBM-resident `FeatureStore` without the ETL pipeline.

Inputs are what `merge_state_actions.py` has in hand at :113 after merging heat metrics and alerts: one row
per (county, date), sorted by county then date, with `HImax_C`, the alert flag, `issued_in_advance`
(NaN = none) and `significance`.  Two things stay on the host: the holiday calendar (the reference asks the
`holidays` package, :184-187; `us_summer_holidays` is the built-in stand-in) and the string handling.
"""
from __future__ import annotations

import ctypes as C
import datetime as dt
from typing import Sequence

import numpy as np
import torch

from . import _lib, schema
from .store import FeatureStore


def us_summer_holidays(years) -> set:
    """US federal holidays between May and September as `datetime.date`s: Memorial Day (last Monday of May),
    Independence Day and its observed weekday, Labor Day (first Monday of September), Juneteenth from 2021."""
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


def holiday_flags(date_days: np.ndarray, calendar=None) -> np.ndarray:
    """uint8 flag per row: the day (days since 1970-01-01) is in `calendar` (default: `us_summer_holidays`)."""
    days = np.asarray(date_days, dtype=np.int64)
    if calendar is None:
        y0 = int(days.min().astype("datetime64[D]").astype("datetime64[Y]").astype(int)) + 1970
        y1 = int(days.max().astype("datetime64[D]").astype("datetime64[Y]").astype(int)) + 1970
        calendar = us_summer_holidays(range(y0, y1 + 1))
    epoch = dt.date(1970, 1, 1)
    hol = np.array(sorted((d - epoch).days for d in calendar), dtype=np.int64)
    return np.isin(days, hol).astype(np.uint8)


class FeatureEngineer:
    """One C context on one GPU for `w2a_engineer_features` (no tables needed)."""

    MAX_ROWS_PER_COUNTY = 200 * 1024 // 32

    def __init__(self, device="cuda:0"):
        self._lib = _lib.lib()  # raises ImportError when the extension is not built: there is no CPU path
        self.device = torch.device(device)
        if self.device.type != "cuda" or not torch.cuda.is_available():
            raise RuntimeError("FeatureEngineer needs a CUDA device; there is no CPU fallback")
        index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", index)
        ctx = C.c_void_p()
        _lib.check(self._lib.w2a_create(index, C.byref(ctx)))
        self.ctx = ctx

    def __del__(self):
        ctx = getattr(self, "ctx", None)
        if ctx:
            try:
                self._lib.w2a_destroy(ctx)
            except Exception:
                pass
            self.ctx = None

    def engineer(self, county_start, date_days, himax_c, alert, issued_in_advance, holiday):
        """Arrays over rows sorted by (county, date); `county_start` = row offsets [S + 1].
        Returns ``(feat float64 [N, 27] cuda, year int32 [N] cuda)``; columns in `schema.REWARD_FEATURES` order."""
        dev = self.device
        start = np.ascontiguousarray(county_start, dtype=np.int64)
        n, S = int(start[-1]), len(start) - 1
        sizes = np.diff(start)
        if (sizes < 0).any() or start[0] != 0:
            raise ValueError("county_start must be non-decreasing row offsets starting at 0")
        if sizes.max(initial=0) > self.MAX_ROWS_PER_COUNTY:
            raise ValueError(f"a county has {int(sizes.max())} rows, more than {self.MAX_ROWS_PER_COUNTY}")
        d = lambda a, t: torch.from_numpy(np.ascontiguousarray(a, dtype=t)).to(dev)  # noqa: E731
        cols = [d(date_days, np.int32), d(himax_c, np.float64), d(alert, np.uint8), d(issued_in_advance, np.float64),
                d(holiday, np.uint8)]
        for c in cols:
            if c.numel() != n:
                raise ValueError(f"every column must have {n} rows")
        t_start = torch.from_numpy(start).to(dev)
        feat = torch.empty((n, schema.N_REWARD_FEATURES), dtype=torch.float64, device=dev)
        year = torch.empty(n, dtype=torch.int32, device=dev)
        work = torch.empty(_lib.FEATURE_WORKSPACE_BYTES, dtype=torch.uint8, device=dev)
        p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(self._lib.w2a_engineer_features(self.ctx, stream, n, S, p(t_start), int(sizes.max(initial=1)),
                                                   *(p(c) for c in cols), p(feat), p(year), p(work)), self.ctx)
        return feat, year


def store_from_raw(fips: Sequence[str], date_days, himax_c, alert, issued_in_advance, significance=None,
                   holiday=None, conf_fips: Sequence[str] = (), conf_zone: Sequence[str] = (), device="cuda:0",
                   max_days: int = schema.MAX_DAYS) -> FeatureStore:
    """Raw rows (sorted by county, then date) -> `FeatureStore`, the features computed on the GPU.

    `fips`: county code per row; `significance`: label or None per row (kept as the store's code column);
    `holiday`: uint8 per row, default `holiday_flags(date_days)`."""
    fips = np.asarray(fips).astype(str)
    days = np.asarray(date_days, dtype=np.int64)
    n = len(fips)
    change = np.flatnonzero(fips[1:] != fips[:-1]) + 1
    start = np.r_[0, change, n].astype(np.int64)
    counties = fips[start[:-1]].tolist()
    if len(set(counties)) != len(counties) or counties != sorted(counties):
        raise ValueError("rows must be sorted by county")
    if (np.diff(days)[np.setdiff1d(np.arange(n - 1), change - 1)] <= 0).any():
        raise ValueError("rows of a county must be sorted by date, one row per day")
    if holiday is None:
        holiday = holiday_flags(days)
    feat_d, year_d = FeatureEngineer(device).engineer(start, days, himax_c, alert, issued_in_advance, holiday)
    feat, year = feat_d.cpu().numpy(), year_d.cpu().numpy()
    years = sorted(int(y) for y in np.unique(year))
    S, Y = len(counties), len(years)
    s_idx = np.repeat(np.arange(S), np.diff(start))
    y_idx = np.searchsorted(np.array(years), year)
    day = feat[:, schema.DATA_COLUMNS.index("dos")].astype(np.int64)
    if day.max(initial=0) >= max_days:
        raise ValueError(f"an episode has {int(day.max()) + 1} days, more than max_days={max_days}")
    n_days = np.zeros((S, Y), dtype=np.uint8)
    np.maximum.at(n_days, (s_idx, y_idx), (day + 1).astype(np.uint8))
    dense = np.zeros((S, Y, max_days, schema.N_REWARD_FEATURES), dtype=np.float64)
    dense[s_idx, y_idx, day] = feat
    sig_code = np.full((S, Y, max_days), np.nan, dtype=np.float32)
    labels: list[str] = []
    if significance is not None:
        sig = np.asarray(significance, dtype=object)
        has = np.array([x is not None and x == x for x in sig])
        labels = sorted({str(x) for x in sig[has]})
        lut = {lab: float(i + 1) for i, lab in enumerate(labels)}
        codes = np.array([lut[str(x)] if h else np.nan for x, h in zip(sig, has)], dtype=np.float32)
        sig_code[s_idx, y_idx, day] = codes
    year_order = [int(y) for y in dict.fromkeys(year.tolist())]  # order of first appearance (env.py:104-105)
    return FeatureStore(fips=counties, years=years, n_days=n_days, feat=dense, sig_code=sig_code, sig_labels=labels,
                        conf_fips=[str(f) for f in conf_fips], conf_zone=[str(z) for z in conf_zone], year_order=year_order)
