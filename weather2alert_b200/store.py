"""Host side of the weather feature store and of the reward-model checkpoints.

`FeatureStore` is the dense form of the three parquet files the reference downloads and
merges in its constructor (`/root/reference/src/weather2alert/env.py:40-57`);
`RewardWeights` is the dense form of ``posterior_samples.safetensors`` + ``config.yaml``
(`env.py:60-85`, written by `/root/reference/reward-training/train.py:121-137`).  Both
know how to pack themselves into the HBM layouts the kernels read (DESIGN.md "Data
layout in HBM"); the packing is pure numpy so it is testable without a GPU.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Sequence

import numpy as np

from . import schema

X_STRIDE = 32  # float64 per day-row record
W_STRIDE = 64  # float32 per coefficient record
X_SIG = 24  # record position of the significance code
X_GATE = 25  # record position of the heat_qi gate column: 0.0 if heat_qi > 0.5 else GATE_OFF
GATE_OFF = -1e300  # times the constant effectiveness coefficient 1.0: sigmoid(z_e) becomes exactly 0
W_OTHER = 56  # first float of the non-table coefficients (lag1, streak, remaining, bias)


def coef_pos_table(u: int, head: int) -> int:
    """Float offset of (table feature `u`, head) in a coefficient record."""
    return 2 * u + head


def coef_pos_other(k: int, head: int) -> int:
    """Float offset of non-table term `k` (0 lag1, 1 streak, 2 remaining, 3 bias)."""
    return W_OTHER + 2 * k + head


#: coefficient slot (0..26 feature, 27 bias) -> float offset of its baseline coefficient
_COEF_POS = np.zeros(schema.N_COEF, dtype=np.int64)
for _u, _slot in enumerate(schema.TABLE_REWARD_SLOTS):
    _COEF_POS[_slot] = coef_pos_table(_u, 0)
_COEF_POS[schema.SLOT_ALERT_LAG1] = coef_pos_other(0, 0)
_COEF_POS[schema.SLOT_ALERT_STREAK] = coef_pos_other(1, 0)
_COEF_POS[schema.SLOT_REMAINING_BUDGET] = coef_pos_other(2, 0)
_COEF_POS[27] = coef_pos_other(3, 0)


@dataclass
class FeatureStore:
    """Dense feature store: one row per (county, year, day)."""

    fips: list[str]  # S counties, data order
    years: list[int]  # Y years, ascending
    n_days: np.ndarray  # [S, Y] uint8, 0 = the reference would raise KeyError (env.py:127)
    feat: np.ndarray  # [S, Y, T, 27] float64, slots 0..26 (bool/int columns as float64)
    sig_code: np.ndarray  # [S, Y, T] float32: NaN = no alert, k+1 = sig_labels[k]
    sig_labels: list[str] = field(default_factory=list)
    conf_fips: list[str] = field(default_factory=list)  # confounders frame, row order
    conf_zone: list[str] = field(default_factory=list)
    #: years in order of first appearance in the data = the reference's default
    #: `valid_years` (env.py:104-105); None = ascending
    year_order: list[int] | None = None

    def __post_init__(self):
        if self.year_order is None:
            self.year_order = [int(y) for y in self.years]
        self.fips_index = {f: i for i, f in enumerate(self.fips)}
        self.year_index = {int(y): i for i, y in enumerate(self.years)}

    @property
    def shape(self):
        return self.feat.shape[:3]  # S, Y, T

    @property
    def budget0(self) -> np.ndarray:
        """Historical ``remaining_budget`` of day 0, the default budget (env.py:169; Q9)."""
        return self.feat[:, :, 0, schema.SLOT_REMAINING_BUDGET].astype(np.int32)

    # ---- packing for the device -----------------------------------------------------
    def pack_x(self) -> np.ndarray:
        """float64 [S*Y*T, 32] day-row records: positions 0..23 the reward features that are
        read from the table (obs slots 0..22 and 26), 24 the significance code with 0.0 for
        "none" (the kernels show it as NaN), 25 the heat_qi gate column, 26..31 zero."""
        S, Y, T = self.shape
        out = np.zeros((S * Y * T, X_STRIDE), dtype=np.float64)
        out[:, :24] = self.feat[..., list(schema.TABLE_REWARD_SLOTS)].reshape(S * Y * T, 24)
        out[:, X_SIG] = np.nan_to_num(self.sig_code.reshape(-1).astype(np.float64), nan=0.0)
        # env.py:218 `effectiveness * (heat_qi > 0.5)` (quirk Q4) as a table column (NaN compares False)
        heat_qi = self.feat[..., schema.SLOT_HEAT_QI].reshape(-1)
        out[:, X_GATE] = np.where(heat_qi > 0.5, 0.0, GATE_OFF)
        return out

    # ---- constructors -----------------------------------------------------------------
    @classmethod
    def from_frames(cls, exogenous, endogenous, confounders, max_days: int = schema.MAX_DAYS):
        """Build from the three reference frames (pandas), merged as at env.py:49-56."""
        import pandas as pd

        merged = pd.merge(exogenous, endogenous, on=["fips", "date"])
        merged["year"] = merged.date.str[:4].astype(int)
        missing = [c for c in schema.DATA_COLUMNS if c not in merged.columns]
        if missing:
            raise KeyError(f"feature store is missing columns {missing}")
        # `merged.loc[(fips, year)]` (env.py:127) yields the rows of a group in file order,
        # so group with a stable sort and keep the order inside each group untouched
        year_order = [int(y) for y in merged.year.unique()]
        merged = merged.sort_values(["fips", "year"], kind="stable")
        fips = sorted(merged.fips.unique().tolist())
        years = sorted(int(y) for y in merged.year.unique())
        S, Y = len(fips), len(years)
        f_idx = merged.fips.map({f: i for i, f in enumerate(fips)}).to_numpy()
        y_idx = merged.year.map({y: i for i, y in enumerate(years)}).to_numpy()
        sy = f_idx * Y + y_idx
        # day index = position inside the (fips, year) group
        starts = np.r_[0, np.flatnonzero(np.diff(sy)) + 1]
        counts = np.diff(np.r_[starts, len(sy)])
        day = np.arange(len(sy)) - np.repeat(starts, counts)
        if counts.max(initial=0) > max_days:
            raise ValueError(f"an episode has {counts.max()} days, more than max_days={max_days}")
        n_days = np.zeros((S, Y), dtype=np.uint8)
        n_days.reshape(-1)[sy[starts]] = counts
        feat = np.zeros((S, Y, max_days, 27), dtype=np.float64)
        numeric = merged[list(schema.REWARD_FEATURES)].to_numpy(dtype=np.float64)
        feat[f_idx, y_idx, day] = numeric
        sig = merged["significance"]
        labels = sorted(str(x) for x in sig.dropna().unique())
        codes = sig.map({lab: float(i + 1) for i, lab in enumerate(labels)}).to_numpy(dtype=np.float32)
        sig_code = np.full((S, Y, max_days), np.nan, dtype=np.float32)
        sig_code[f_idx, y_idx, day] = codes
        return cls(
            fips=fips,
            years=years,
            n_days=n_days,
            feat=feat,
            sig_code=sig_code,
            sig_labels=labels,
            conf_fips=[str(f) for f in confounders.fips.tolist()],
            conf_zone=[str(z) for z in confounders.ba_zone.tolist()],
            year_order=year_order,
        )

    @classmethod
    def from_parquet_dir(cls, path: str):
        """Read ``{exogenous_states,endogenous_states_actions,confounders}.parquet`` (env.py:40-57)."""
        import pandas as pd

        frames = {}
        for name in ("exogenous_states", "endogenous_states_actions", "confounders"):
            p = os.path.join(path, name + ".parquet")
            if not os.path.exists(p):
                raise FileNotFoundError(p)
            frames[name] = pd.read_parquet(p)
        return cls.from_frames(
            frames["exogenous_states"], frames["endogenous_states_actions"], frames["confounders"]
        )


@dataclass
class RewardWeights:
    """Posterior coefficient draws: ``coef[p, s, head, k]``, head 0 baseline / 1 effectiveness,
    k = feature slot 0..26 or 27 = bias."""

    fips_list: list[str]
    coef: np.ndarray  # [P, S, 2, 28] float32

    @property
    def n_samples(self) -> int:
        return self.coef.shape[0]

    def pack(self) -> np.ndarray:
        """float32 [S*P, 64] coefficient records: record ``col*P + p`` = 24 table features x
        (b,e), the gate column's (0,1) at floats 50,51 inside 8 otherwise-zero floats, then the four
        non-table terms x (b,e) (see `coef_pos_*`)."""
        P, S = self.coef.shape[:2]
        out = np.zeros((S, P, W_STRIDE), dtype=np.float32)
        c = np.transpose(self.coef, (1, 0, 2, 3))  # [S, P, 2, 28]
        out[:, :, _COEF_POS] = c[:, :, 0, :]
        out[:, :, _COEF_POS + 1] = c[:, :, 1, :]
        out[:, :, 2 * X_GATE + 1] = 1.0  # effectiveness coefficient of the gate column
        return out.reshape(S * P, W_STRIDE)

    @classmethod
    def from_tensors(cls, tensors: dict, fips_list: Sequence) -> "RewardWeights":
        """`tensors`: the 56 ``baseline_*`` / ``effectiveness_*`` arrays of shape [P, 1, S]."""
        names = list(schema.REWARD_FEATURES) + ["bias"]
        first = np.asarray(tensors["baseline_bias"])
        P, _, S = first.shape
        coef = np.zeros((P, S, 2, schema.N_COEF), dtype=np.float32)
        for h, head in enumerate(("baseline", "effectiveness")):
            for k, name in enumerate(names):
                coef[:, :, h, k] = np.asarray(tensors[f"{head}_{name}"], dtype=np.float32)[:, 0, :]
        return cls(fips_list=[str(x) for x in fips_list], coef=coef)

    @classmethod
    def from_dir(cls, path: str) -> "RewardWeights":
        """Read ``posterior_samples.safetensors`` + ``config.yaml`` (env.py:60-85)."""
        import yaml
        from safetensors.numpy import load_file

        tensors = load_file(os.path.join(path, "posterior_samples.safetensors"))
        with open(os.path.join(path, "config.yaml")) as fh:
            config = yaml.safe_load(fh)
        # the reference raises KeyError when the checkpoint has no county list (env.py:75)
        return cls.from_tensors(tensors, config["fips_list"])

    def to_tensors(self) -> dict:
        """Inverse of `from_tensors` (the reference checkpoint layout)."""
        names = list(schema.REWARD_FEATURES) + ["bias"]
        out = {}
        for h, head in enumerate(("baseline", "effectiveness")):
            for k, name in enumerate(names):
                out[f"{head}_{name}"] = np.ascontiguousarray(self.coef[:, None, :, h, k])
        return out
