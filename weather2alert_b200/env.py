"""`HeatAlertEnv` -- drop-in for the reference's single-env Gymnasium class, running on a B200.

Same constructor, `reset` and `step` signatures, RNG draw order, budget rules and info
dict as `/root/reference/src/weather2alert/env.py:17-262`; the arithmetic of every
`reset`/`step` happens in the sm_100a kernels behind `libw2a.so` (batch of one), reached
through the host-buffer C entry points.  Differences, all deliberate:

* observations are ``float32[29]`` (the reference returns an ``object`` array with a
  string/NaN ``significance`` slot, quirk Q10): slot 27 is NaN or ``1 + index`` into
  `sig_labels`; `observation_space.shape` is ``(29,)``, the real length, not the
  reference's ``(33,)``;
* `seed=` is accepted by the constructor (README form ``HeatAlertEnv(seed=1234)``) and
  used for resets that do not pass their own seed;
* data and weights can be handed over as objects (`store=`, `weights=RewardWeights`),
  which is how tests and benchmarks run without the Hugging Face hub.
"""
from __future__ import annotations

from typing import Literal

import numpy as np

from . import schema
from ._gym import Env, spaces
from .core import HeatAlertCore
from .loading import load_store, load_weights
from .sampling import EpisodeSampler
from .store import FeatureStore, RewardWeights


class HeatAlertEnv(Env):
    """Class to simulate the environment for the online RL agent (GPU-backed)."""

    def __init__(
        self,
        weights: str | RewardWeights = "nn_full_medicare_all",
        years: list | None = None,
        fips_list: list | None = None,  # ignored, as in the reference (env.py:75 overwrites it)
        similar_climate_counties: bool = False,
        budget: int | None = None,
        data_dir: str | None = None,
        split: str = "65k",
        *,
        seed: int | None = None,
        store: FeatureStore | None = None,
        device: str = "cuda:0",
    ):
        super().__init__()
        self.similar_climate_counties = similar_climate_counties
        self.budget = budget
        self._default_seed = seed
        self.store = store if store is not None else load_store(data_dir, split)
        self.weights = weights if isinstance(weights, RewardWeights) else load_weights(weights, data_dir)
        self.fips_list = list(self.weights.fips_list)
        self.n_samples = self.weights.n_samples
        self.sampler = EpisodeSampler(self.store, self.weights, years)
        self.valid_years = self.sampler.valid_years
        self.sig_labels = list(self.store.sig_labels)
        self.observation_space = spaces.Box(low=-np.inf, high=np.inf, shape=(schema.OBS_DIM,), dtype=np.float32)
        self.action_space = spaces.Discrete(2)  # alert or no alert
        self.feat_names = list(schema.OBS_NAMES)

        # raises without GPU / built extension; a batch of one does not pay for the whole-summer tables (3 GB)
        self.core = HeatAlertCore(self.store, self.weights, device, precompute=False)
        self.core.host_envs_create(1)
        self._obs = np.zeros((1, schema.OBS_DIM), np.float32)
        self._reward = np.zeros(1, np.float32)
        self._term = np.zeros(1, np.uint8)
        self._rem = np.zeros(1, np.int32)
        self._atb = np.zeros(1, np.uint8)
        self._act = np.zeros(1, np.uint8)

    # ---- reset (env.py:133-184) ----------------------------------------------------------
    def reset(
        self,
        location: str | None = None,
        similar_climate_counties: bool | None = None,
        seed: int | None = None,
        budget: int | None = None,
        sample_budget: bool = False,
        sample_budget_type: Literal["less_than", "centered"] = "less_than",
        options: dict | None = None,  # gymnasium's keyword; unused
    ):
        if seed is None:
            seed = self._default_seed
            self._default_seed = None  # like gymnasium: a ctor seed applies to the first reset only
        if seed is None:
            seed = np.random.randint(0, 10000)  # env.py:143-144
        self.rng = np.random.default_rng(seed)  # env.py:145
        if similar_climate_counties is None:
            similar_climate_counties = self.similar_climate_counties

        d = self.sampler.draw_reference(self.rng, location, similar_climate_counties)
        self.location = d["info_location"]
        self.location_index = d["location_index"]
        self.ep_index = d["location"] + "_" + str(d["year"])  # env.py:156
        self.n_days = int(self.store.n_days[d["s"], d["y"]])
        self.coef_index = d["coef_index"]
        self.attempted_alert_buffer = []
        self.actual_alert_buffer = []
        self.alert_streak = 0
        self.t = 0

        if self.budget is None:  # env.py:167-170; sticky from now on (Q8)
            self.budget = int(self.store.budget0[d["s"], d["y"]]) if budget is None else budget
        if sample_budget:  # env.py:172-177
            b = self.budget
            if sample_budget_type == "less_than":
                self.budget = int(self.rng.integers(0, b + 1))
            elif sample_budget_type == "centered":
                self.budget = int(self.rng.integers(0.5 * b, 1.5 * b + 1))
        self.remaining_budget = self.budget
        self.at_budget = False

        i32 = lambda v: np.array([v], dtype=np.int32)  # noqa: E731
        self.core.reset_host(i32(d["s"]), i32(d["y"]), i32(self.location_index), i32(self.coef_index),
                             self._obs, budget=i32(self.budget), remaining=self._rem)
        self.observation = self._obs[0].copy()
        return self.observation, self._get_info()

    def _get_info(self) -> dict:  # env.py:228-236
        return {
            "episode_index": self.ep_index,
            "remaining_budget": self.remaining_budget,
            "at_budget": self.at_budget,
            "feature_names": self.feat_names,
            "location": self.location,
            "location_index": self.location_index,
        }

    # ---- step (env.py:238-262) -------------------------------------------------------------
    def step(self, action: int):
        action = int(action)
        self._act[0] = 1 if action else 0
        self.core.step_host(self._act, self._obs, self._reward, self._term, self._rem, self._atb)
        self.at_budget = bool(self._atb[0])
        self.remaining_budget = int(self._rem[0])
        actual_action = 0 if (action == 1 and self.at_budget) else action
        self.attempted_alert_buffer.append(action)
        self.actual_alert_buffer.append(actual_action)
        done = bool(self._term[0])
        if not done:
            self.t += 1
            self.alert_streak = self.alert_streak + 1 if actual_action else 0
        self.observation = self._obs[0].copy()
        return self.observation, float(self._reward[0]), done, False, self._get_info()

    def close(self):
        self.core = None
