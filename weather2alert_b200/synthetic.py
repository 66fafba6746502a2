"""Synthetic feature stores and checkpoints with the reference's schema.

The real data and weights live on the Hugging Face hub (`env.py:40-47,60-67`) and are
not reachable here, so tests, the oracle harness and `bench.py`
all run on synthetic fixtures.  Value ranges follow the ETL that produces the real
files (`/root/reference/data-processing/merge_state_actions.py:119-210`; SURVEY.md
appendix C); the alert-history columns are *derived* from the sampled historical alerts
exactly as the ETL derives them, so the default budget (quirk Q9) is consistent.
"""
from __future__ import annotations

import datetime as _dt
from typing import Sequence

import numpy as np

from . import schema
from .store import FeatureStore, RewardWeights

_ZONES = ("Cold", "Mixed-Humid", "Hot-Humid", "Hot-Dry", "Marine", "Mixed-Dry", "Very Cold")
_ZONE_P = (0.40, 0.35, 0.15, 0.05, 0.03, 0.01, 0.01)
_STATES = (
    "01 04 05 06 08 09 10 12 13 16 17 18 19 20 21 22 24 25 26 27 28 29 30 31 32 33 34 35 36 37 "
    "38 39 40 41 42 44 45 46 47 48 49 50 51 53 54 55 56"
).split()


def synthetic_counties(n: int, seed: int = 0, must_include: Sequence[str] = ("06037",)):
    """`n` distinct 5-digit FIPS codes (sorted) and a climate zone for each."""
    rng = np.random.default_rng(seed)
    codes = set(must_include)
    while len(codes) < n:
        st = _STATES[int(rng.integers(len(_STATES)))]
        codes.add(f"{st}{2 * int(rng.integers(0, 420)) + 1:03d}")
    fips = sorted(codes)[:n] if len(codes) > n else sorted(codes)
    zones = [str(_ZONES[i]) for i in rng.choice(len(_ZONES), size=len(fips), p=_ZONE_P)]
    return fips, zones


def make_store(
    fips: Sequence[str],
    years: Sequence[int],
    seed: int = 0,
    zones: Sequence[str] | None = None,
    n_days: int | np.ndarray = schema.MAX_DAYS,
    p_alert: float = 0.05,
    extra_conf: Sequence[tuple[str, str]] = (),
) -> FeatureStore:
    """Random feature store: S counties x Y years x T days, float64 features."""
    rng = np.random.default_rng(seed)
    S, Y, T = len(fips), len(years), schema.MAX_DAYS
    feat = np.zeros((S, Y, T, 27), dtype=np.float64)
    col = {name: i for i, name in enumerate(schema.REWARD_FEATURES)}

    hq = rng.random((S, Y, T))
    hi = 0.6 + 0.5 * rng.random((S, Y, T))
    feat[..., col["heat_qi"]] = hq
    feat[..., col["heat_qi_above_25"]] = hq * (hq > 0.25)
    feat[..., col["heat_qi_above_75"]] = hq * (hq > 0.75)
    feat[..., col["hi_max"]] = hi
    # real data compares 0.01*degF against 25/75, so these two are identically 0 there
    # (appendix C); keep a few non-zeros so that their coefficients are exercised
    feat[..., col["hi_max_above_25"]] = hi * (rng.random((S, Y, T)) < 0.1)
    feat[..., col["hi_max_above_75"]] = hi * (rng.random((S, Y, T)) < 0.02)
    feat[..., col["hi_max*heat_qi"]] = hq * hi
    feat[..., col["hi_max_above_25*heat_qi"]] = feat[..., col["heat_qi_above_25"]] * hi
    feat[..., col["hi_max_above_75*heat_qi"]] = feat[..., col["heat_qi_above_75"]] * hi
    c3 = np.cumsum(hq, axis=2)
    q3 = hq.copy()
    q7 = hq.copy()
    for t in range(T):
        lo3, lo7 = max(0, t - 2), max(0, t - 6)
        q3[..., t] = (c3[..., t] - (c3[..., lo3 - 1] if lo3 > 0 else 0)) / (t - lo3 + 1)
        q7[..., t] = (c3[..., t] - (c3[..., lo7 - 1] if lo7 > 0 else 0)) / (t - lo7 + 1)
    feat[..., col["heat_qi_3d"]] = q3
    feat[..., col["heat_qi_7d"]] = q7
    feat[..., col["excess_heat_3d"]] = np.clip(hq - q3, 0, None)
    feat[..., col["excess_heat_7d"]] = np.clip(hq - q7, 0, None)
    feat[..., col["excess_heat_3d*heat_qi"]] = feat[..., col["excess_heat_3d"]] * hq
    feat[..., col["excess_heat_7d*heat_qi"]] = feat[..., col["excess_heat_7d"]] * hq
    feat[..., col["weekend"]] = rng.random((S, Y, T)) < 0.28
    feat[..., col["holiday"]] = rng.random((S, Y, T)) < 0.02
    dos = np.arange(T, dtype=np.float64)
    feat[..., col["dos"]] = dos
    u = dos / (T - 1)
    basis = np.stack([3 * u * (1 - u) ** 2, 3 * u**2 * (1 - u), u**3], axis=-1)
    basis = (basis - basis.mean(0)) / basis.std(0)
    for i in range(3):
        feat[..., col[f"bspline_dos_{i}"]] = basis[:, i]

    alert = rng.random((S, Y, T)) < p_alert
    ca = np.cumsum(alert, axis=2)
    feat[..., col["alert"]] = alert
    w14 = ca.copy()
    w14[..., 14:] -= ca[..., :-14]
    feat[..., col["alerts_2wks"]] = w14
    feat[..., 1:, col["alert_lag1"]] = alert[..., :-1]
    streak = np.zeros((S, Y, T))
    run = np.zeros((S, Y))
    for t in range(T):
        run = np.where(alert[..., t], run + 1, 0)
        streak[..., t] = run
    feat[..., col["alert_streak"]] = streak
    feat[..., col["remaining_budget"]] = ca[..., -1:] - ca
    feat[..., col["issued_in_advance"]] = np.where(alert, np.round(3 * rng.random((S, Y, T)), 2), 0.0)
    labels = ["A", "W", "Y"]
    sig = np.where(alert, rng.integers(1, 4, (S, Y, T)).astype(np.float32), np.float32("nan"))

    nd = np.broadcast_to(np.asarray(n_days, dtype=np.uint8), (S, Y)).copy()
    for s in range(S):
        for y in range(Y):
            feat[s, y, nd[s, y] :] = 0.0
            sig[s, y, nd[s, y] :] = np.nan
    if zones is None:
        zones = [str(_ZONES[i]) for i in rng.choice(len(_ZONES), size=S, p=_ZONE_P)]
    conf_fips = list(fips) + [f for f, _ in extra_conf]
    conf_zone = list(zones) + [z for _, z in extra_conf]
    order = np.argsort(np.array(conf_fips), kind="stable")
    return FeatureStore(
        fips=list(fips),
        years=[int(y) for y in years],
        n_days=nd,
        feat=feat,
        sig_code=sig.astype(np.float32),
        sig_labels=labels,
        conf_fips=[conf_fips[i] for i in order],
        conf_zone=[conf_zone[i] for i in order],
    )


def make_weights(fips_list: Sequence[str], n_samples: int = 100, seed: int = 0, scale: float = 0.3) -> RewardWeights:
    """``scale * N(0,1)`` float32 coefficient draws in the checkpoint layout."""
    rng = np.random.default_rng(seed)
    coef = (scale * rng.standard_normal((n_samples, len(fips_list), 2, schema.N_COEF))).astype(np.float32)
    # keep the logits in a realistic range: `dos` runs to 152 and `remaining_budget` into the tens
    coef[..., 17] *= 0.02
    coef[..., schema.SLOT_REMAINING_BUDGET] *= 0.1
    return RewardWeights(fips_list=[str(f) for f in fips_list], coef=coef)


def store_to_frames(store: FeatureStore):
    """The three pandas frames the reference reads, with the dtypes the ETL writes
    (`merge_state_actions.py:228-281`)."""
    import pandas as pd

    S, Y, T = store.shape
    rows = [(s, y, t) for s in range(S) for y in range(Y) for t in range(int(store.n_days[s, y]))]
    si, yi, ti = (np.array(v) for v in zip(*rows))
    fips = np.array(store.fips, dtype=object)[si]
    may1 = {y: _dt.date(int(y), 5, 1) for y in store.years}
    dates = np.array(
        [(may1[store.years[y]] + _dt.timedelta(days=int(t))).isoformat() for y, t in zip(yi, ti)], dtype=object
    )
    vals = store.feat[si, yi, ti]
    exo = pd.DataFrame({name: vals[:, i] for i, name in enumerate(schema.EXOGENOUS)})
    for name in ("weekend", "holiday", "dos"):
        exo[name] = exo[name].astype(np.int64)
    exo["fips"] = fips
    exo["date"] = dates
    endo = pd.DataFrame({"fips": fips, "date": dates})
    off = len(schema.EXOGENOUS)
    endo["alert"] = vals[:, off + 0].astype(bool)
    endo["alerts_2wks"] = vals[:, off + 1]
    endo["alert_lag1"] = vals[:, off + 2].astype(np.int64)
    endo["alert_streak"] = vals[:, off + 3].astype(np.int64)
    endo["remaining_budget"] = vals[:, off + 4].astype(np.int64)
    endo["issued_in_advance"] = vals[:, off + 5]
    codes = store.sig_code[si, yi, ti]
    sig = np.array([None if np.isnan(c) else store.sig_labels[int(c) - 1] for c in codes], dtype=object)
    endo["significance"] = pd.Series(sig, dtype="object")
    conf = pd.DataFrame({"fips": list(store.conf_fips), "ba_zone": list(store.conf_zone)})
    return exo, endo, conf


def write_reference_files(store: FeatureStore, weights: RewardWeights, out_dir: str) -> None:
    """Write the five files `HeatAlertEnv.__init__` fetches (`env.py:40,60`) into one flat dir."""
    import os

    import yaml
    from safetensors.numpy import save_file

    os.makedirs(out_dir, exist_ok=True)
    exo, endo, conf = store_to_frames(store)
    exo.to_parquet(os.path.join(out_dir, "exogenous_states.parquet"))
    endo.to_parquet(os.path.join(out_dir, "endogenous_states_actions.parquet"))
    conf.to_parquet(os.path.join(out_dir, "confounders.parquet"))
    save_file(weights.to_tensors(), os.path.join(out_dir, "posterior_samples.safetensors"))
    with open(os.path.join(out_dir, "config.yaml"), "w") as fh:
        yaml.safe_dump({"name": "synthetic", "fips_list": list(weights.fips_list)}, fh)
