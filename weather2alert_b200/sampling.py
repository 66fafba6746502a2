"""Episode sampling: the host-side half of `reset` (index streams for the reset kernel).

The reference seeds a fresh PCG64 in every `reset` and draws, in this order
(`/root/reference/src/weather2alert/env.py:143-177`):
location (only if not given), index into the similar-county list (only when augmenting),
year, posterior sample, budget (only with `sample_budget`).  `draw_reference` makes the
same numpy calls in the same order so that a seeded reset reproduces the reference;
`draw_bulk` draws whole batches from one Generator for the vector env.
"""
from __future__ import annotations

from typing import Sequence

import numpy as np

from .datautils import ClimateGroups
from .store import FeatureStore, RewardWeights


class EpisodeSampler:
    def __init__(self, store: FeatureStore, weights: RewardWeights, years: Sequence[int] | None = None):
        self.store = store
        self.fips_list = list(weights.fips_list)
        self.n_samples = weights.n_samples
        # env.py:32,104-105: the ctor's `years` or every year present, in order of appearance
        self.valid_years = list(years) if years is not None else list(store.year_order)
        self.groups = ClimateGroups(store.conf_fips, store.conf_zone, self.fips_list)
        self.list_to_data = np.array([store.fips_index.get(f, -1) for f in self.fips_list], dtype=np.int64)
        self.n_similar = self.groups.counts_for(self.fips_list)
        # CSR view of the similar-county lists in fips_list indices, for the "intended" augmentation:
        # county i of the weights -> members[similar_start[i] : similar_start[i] + n_similar[i]]
        lut = {f: i for i, f in enumerate(self.fips_list)}
        self.similar_members = np.array([lut[f] for m in self.groups.members for f in m], dtype=np.int64)
        self.similar_start = np.array([self.groups.offsets[self.groups.group_of[f]] if f in self.groups.group_of else 0
                                       for f in self.fips_list], dtype=np.int64)
        self._valid_year_idx = np.array([store.year_index.get(int(y), -1) for y in self.valid_years], dtype=np.int64)

    # ---- one env, reference draw order ----------------------------------------------------
    def episode(self, location: str, year) -> tuple[int, int]:
        """(county, year) -> indices into the feature store; KeyError like `merged.loc` (env.py:127)."""
        s = self.store.fips_index.get(str(location))
        y = self.store.year_index.get(int(year))
        if s is None or y is None or self.store.n_days[s, y] == 0:
            raise KeyError((location, year))
        return s, y

    def draw_reference(self, rng: np.random.Generator, location, augment: bool):
        """Returns dict(location, info_location, location_index, year, s, y, coef_index)."""
        if location is None:
            location = rng.choice(self.fips_list)  # env.py:151-152
        location = str(location)
        if augment:
            similar = self.groups.similar(location)  # env.py:115-116
            location_index = int(rng.choice(range(len(similar))))  # env.py:117 (Q6)
            info_location = similar[location_index]  # env.py:118
        else:
            info_location = location
            location_index = self.fips_list.index(location)  # env.py:121 -> ValueError if unknown
        year = rng.choice(self.valid_years)  # env.py:125
        s, y = self.episode(location, year)  # ORIGINAL location (Q7), env.py:127
        coef_index = int(rng.integers(0, self.n_samples))  # env.py:160
        return dict(location=location, info_location=info_location, location_index=location_index,
                    year=int(year), s=s, y=y, coef_index=coef_index)

    # ---- many envs, one Generator -----------------------------------------------------------
    def draw_bulk(self, rng: np.random.Generator, n: int, location=None, augment=False):
        """Vectorised draws (same distributions, one stream for the whole batch).

        `augment`: False, True (the reference's behaviour: only the coefficient column changes, to the
        index inside the similar-county list -- quirks Q6/Q7), or "intended" (NOT the reference: the
        episode moves to a county drawn from the similar list -- its weather, its own coefficient column).

        `location`: None (uniform over the weights' county list), a FIPS string, or a
        sequence of n FIPS strings / fips_list indices.  Returns int32 arrays
        `loc` (feature-store county), `year`, `coef_col`, `coef_idx` and `list_idx`
        (county as index into fips_list, for `info`)."""
        if location is None:
            list_idx = rng.integers(0, len(self.fips_list), size=n)
        elif isinstance(location, str):
            list_idx = np.full(n, self.fips_list.index(location), dtype=np.int64)
        else:
            location = np.asarray(location)
            if location.dtype.kind in "iu":
                list_idx = location.astype(np.int64)
            else:
                lut = {f: i for i, f in enumerate(self.fips_list)}
                list_idx = np.array([lut[str(f)] for f in location], dtype=np.int64)  # KeyError if unknown
            if list_idx.shape != (n,):
                raise ValueError(f"location must have {n} entries")
        loc = self.list_to_data[list_idx]
        if (loc < 0).any():
            raise KeyError(self.fips_list[int(list_idx[np.argmax(loc < 0)])])
        if augment:
            n_c = self.n_similar[list_idx]
            if (n_c <= 0).any():
                raise KeyError(self.fips_list[int(list_idx[np.argmax(n_c <= 0)])])
            coef_col = rng.integers(0, n_c)  # index inside the similar list (Q6)
            if augment == "intended":  # the similar county itself: its weather and its own column
                list_idx = self.similar_members[self.similar_start[list_idx] + coef_col]
                coef_col = list_idx
                loc = self.list_to_data[list_idx]
                if (loc < 0).any():
                    raise KeyError(self.fips_list[int(list_idx[np.argmax(loc < 0)])])
        else:
            coef_col = list_idx
        year = self._valid_year_idx[rng.integers(0, len(self.valid_years), size=n)]
        if (year < 0).any() or (self.store.n_days[loc, np.maximum(year, 0)] == 0).any():
            bad = int(np.argmax((year < 0) | (self.store.n_days[loc, np.maximum(year, 0)] == 0)))
            raise KeyError((self.fips_list[int(list_idx[bad])], int(year[bad])))
        coef_idx = rng.integers(0, self.n_samples, size=n)
        as32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)  # noqa: E731
        return dict(loc=as32(loc), year=as32(year), coef_col=as32(coef_col), coef_idx=as32(coef_idx),
                    list_idx=as32(list_idx))
