"""`HeatAlertVectorEnv` -- N independent HeatAlertEnv episodes stepped by one kernel launch.

Follows the `gymnasium.vector.VectorEnv` surface (`num_envs`, `single_observation_space`,
`single_action_space`, `reset(seed=, options=)`, `step(actions)`), but every array it
returns is a torch CUDA tensor: observations ``float32[N,29]``, rewards ``float32[N]``,
terminations / truncations ``bool[N]``, and an `infos` dict of tensors.  The reference
has no vector wrapper (users wrap `HeatAlertEnv` with Gymnasium/SB3);
per-env semantics are exactly those of `/root/reference/src/weather2alert/env.py:133-262`.

Episode sampling (`env.py:143-177`: county, similar-county index, year, posterior sample,
budget) is drawn on the device from one seeded torch Generator, so a reset costs a few
tiny launches and never stalls the stream; pass ``options={"index_streams": ...}`` to
supply host-sampled streams instead (that is how trajectories are reproduced against the
reference, whose own resets seed one PCG64 per env).

Autoreset is SAME_STEP: an env that terminates is reset inside the same `step` call (the
returned observation is the first one of its next episode; the reference's terminal
observation is just the previous one, quirk Q2); the terminal step's `remaining_budget` /
`at_budget` are kept in ``infos["final_info"]``.  Termination is ``t >= n_days - 1``,
a function of the sampled episode only, so when all summers have the same length the
host knows the schedule and autoreset needs no device->host synchronisation.

`infos` carries, per env: `remaining_budget`, `at_budget`, `location_index` (the reference's
`info["location_index"]`, env.py:117,121,234: the coefficient column -- the index inside the
similar-county list when augmenting), `location_list_index` (index of `info["location"]` in
`fips_list`) as device tensors, and `episode_index` / `location` as lazily formatted string
sequences (env.py:156,228-236; formatting them synchronises with the device).

Subclasses `gymnasium.vector.VectorEnv` when gymnasium is importable (`_gym.HAVE_GYMNASIUM`).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, schema
from ._gym import VectorEnv, spaces
from .core import HeatAlertCore
from .sampling import EpisodeSampler
from .store import FeatureStore, RewardWeights


class LazyStrings:
    """Per-env strings of `infos` (`episode_index`, `location`; env.py:156,228-236), formatted on the host only
    when somebody looks: indexing, iterating or `numpy()` copies the index tensors back (one synchronisation)."""

    def __init__(self, n, make):
        self._n, self._make, self._cache = n, make, None

    def numpy(self) -> np.ndarray:
        if self._cache is None:
            self._cache = self._make()
        return self._cache

    def __len__(self):
        return self._n

    def __getitem__(self, i):
        return self.numpy()[i]

    def __iter__(self):
        return iter(self.numpy())

    def __repr__(self):
        return f"LazyStrings(n={self._n})"


class HeatAlertVectorEnv(VectorEnv):
    metadata = {"autoreset_mode": "same_step"}

    def __init__(
        self,
        num_envs: int,
        weights: RewardWeights | str = "nn_full_medicare_all",
        store: FeatureStore | None = None,
        years: list | None = None,
        similar_climate_counties: bool = False,
        budget: int | None = None,
        data_dir: str | None = None,
        split: str = "65k",
        device: str = "cuda:0",
        seed: int | None = None,
        autoreset: bool = True,
        core: HeatAlertCore | None = None,
        step_variant: int = _lib.STEP_AUTO,
        track_stats: bool = False,
        sort_episodes: bool | None = None,
    ):
        if core is None:
            if store is None or not isinstance(weights, RewardWeights):
                from .loading import load_store, load_weights

                store = store if store is not None else load_store(data_dir, split)
                weights = weights if isinstance(weights, RewardWeights) else load_weights(weights, data_dir)
            core = HeatAlertCore(store, weights, device)
        self.core = core
        self.store, self.weights = core.store, core.weights
        self.device = core.device
        self.num_envs = int(num_envs)
        self.similar_climate_counties = similar_climate_counties
        self.budget = budget
        self.autoreset = autoreset
        self.step_variant = step_variant
        # Assign the sampled episodes to env slots in (county, coefficient column, sample, year)
        # order.  The draws are i.i.d., so which slot gets which episode is immaterial, but
        # neighbouring envs then share table lines (+20..25 % per step, +35 % per rollout on a B200).
        # Default (None): on, for device-sampled resets of the WHOLE batch only -- caller-supplied
        # streams (`index_streams`, `reference_rng`) fix the slot order, and a masked autoreset (ragged
        # summers) keeps its draws unsorted, because sorting a subset would tie slots to county ranges
        # and over-sample short-summer counties.
        self.sort_episodes = True if sort_episodes is None else bool(sort_episodes)
        self.sampler = EpisodeSampler(self.store, self.weights, years)
        self.fips_list = self.sampler.fips_list
        self.single_observation_space = spaces.Box(-np.inf, np.inf, (schema.OBS_DIM,), np.float32)
        self.single_action_space = spaces.Discrete(2)
        self.observation_space = spaces.Box(-np.inf, np.inf, (self.num_envs, schema.OBS_DIM), np.float32)
        self.action_space = self.single_action_space
        self.feature_names = list(schema.OBS_NAMES)

        n, dev = self.num_envs, self.device
        self._gen = torch.Generator(device=dev)
        self._gen.manual_seed(int(seed) if seed is not None else int(np.random.randint(0, 2**31 - 1)))
        self._host_rng = np.random.default_rng(seed)
        # device copies of the sampler's integer tables
        sp = self.sampler
        self._list_to_data = torch.from_numpy(sp.list_to_data).to(dev)
        self._n_similar = torch.from_numpy(sp.n_similar.astype(np.int64)).to(dev)
        self._similar_start = torch.from_numpy(sp.similar_start).to(dev)
        self._similar_members = torch.from_numpy(sp.similar_members).to(dev)
        self._valid_year_idx = torch.from_numpy(sp._valid_year_idx).to(dev)
        nd = self.store.n_days
        ok_c = sp.list_to_data >= 0
        ok_y = sp._valid_year_idx >= 0
        sub = nd[np.ix_(sp.list_to_data[ok_c], sp._valid_year_idx[ok_y])]
        # every (county of the weights, valid year) exists -> device draws cannot hit a missing episode
        self._all_episodes_exist = bool(ok_c.all() and ok_y.all() and (sub > 0).all())
        self._uniform_days = int(sub.flat[0]) if sub.size and (sub == sub.flat[0]).all() else None
        # every county of the weights has a non-empty similar-county list -> augmented draws cannot fail either
        self._augment_safe = bool((sp.n_similar > 0).all())
        self._n_days_dev = core.n_days  # uint8 [S*Y]

        self._state = core.alloc_state(n)
        self._obs = torch.zeros((n, schema.OBS_DIM), dtype=torch.float32, device=dev)
        self._reward = torch.zeros(n, dtype=torch.float32, device=dev)
        self._term = torch.zeros(n, dtype=torch.bool, device=dev)
        self._trunc = torch.zeros(n, dtype=torch.bool, device=dev)
        self._rem = torch.zeros(n, dtype=torch.int32, device=dev)
        self._atb = torch.zeros(n, dtype=torch.bool, device=dev)
        self._bound = {}
        self._canonical = False  # every env's coefficient column is its own county (set by each draw)
        self._slots_sorted = False  # the env slots are in table order (last full reset was locality-sorted): a hint for the step
        # rollout statistics a la AlertLoggingCallback (callbacks.py:63-74); `stats.reduce()` all-reduces them
        self.stats = None
        if track_stats:
            from .distributed import RolloutStats

            self.stats = RolloutStats(n, dev)
        self._idx = {}        # index streams of the most recent draw (device int32)
        # per env, for `infos`: the episode (store county, store year), the coefficient column, and the
        # fips_list index of the county reported as info["location"]
        self._ep_loc = torch.zeros(n, dtype=torch.int32, device=dev)
        self._ep_year = torch.zeros(n, dtype=torch.int32, device=dev)
        self._coef_col = torch.zeros(n, dtype=torch.int32, device=dev)
        self._list_idx = torch.zeros(n, dtype=torch.int32, device=dev)
        self._final_info = None
        self._term_at = None  # device int64[n]: step number at which each env terminates
        self._step_no = 0
        self._next_term = None
        self._reset_opts = {}

    # ---- sampling -------------------------------------------------------------------------------
    def _randint(self, high, n):
        return torch.randint(0, int(high), (n,), generator=self._gen, device=self.device)

    def _draw(self, sort: bool = True):
        """Index streams for all N envs as device int32 tensors (`list_idx`: the county reported as
        info["location"], as fips_list index)."""
        n, dev, o = self.num_envs, self.device, self._reset_opts
        sp = self.sampler
        if "index_streams" in o:
            src = o["index_streams"]
            s = {k: torch.as_tensor(np.ascontiguousarray(src[k], dtype=np.int32)).to(dev)
                 for k in ("loc", "year", "coef_col", "coef_idx")}
            if "budget" in src:
                s["budget"] = torch.as_tensor(np.ascontiguousarray(src["budget"], dtype=np.int32)).to(dev)
            s["list_idx"] = (torch.as_tensor(np.ascontiguousarray(src["list_idx"], dtype=np.int32)).to(dev)
                             if "list_idx" in src else s["coef_col"])
            for key, hi in (("loc", self.core.S), ("year", self.core.Y), ("coef_col", self.core.Sw),
                            ("coef_idx", self.core.P)):
                a = np.asarray(src[key])
                if a.shape != (n,) or (a < 0).any() or (a >= hi).any():
                    raise IndexError(f"index_streams[{key!r}] must hold {n} values in [0, {hi})")
            nd = self.store.n_days[np.asarray(src["loc"]), np.asarray(src["year"])]
            if (nd == 0).any():
                bad = int(np.argmax(nd == 0))
                raise KeyError((self.store.fips[int(src["loc"][bad])], self.store.years[int(src["year"][bad])]))
            # every env on its own county's column? then the tabulated logits apply (checked on the host streams)
            self._canonical = bool((sp.list_to_data[np.asarray(src["coef_col"])] == np.asarray(src["loc"])).all())
            return s
        augment = o.get("similar_climate_counties", self.similar_climate_counties)
        self._canonical = (not augment) or augment == "intended"
        location = o.get("location")
        if o.get("reference_rng", False):
            # what gymnasium.vector.SyncVectorEnv([HeatAlertEnv]*N).reset(seed=s) does: env i seeds its
            # own PCG64 with s + i and draws in the reference order (env.py:143-177).  Host loop: for
            # parity and small batches.
            if augment == "intended":
                raise ValueError("reference_rng reproduces the reference's draws; 'intended' augmentation is not one of them")
            base_seed = o["reference_seed"]
            locs = location if isinstance(location, (list, tuple, np.ndarray)) else [location] * n
            rows, budgets = [], []
            for i in range(n):
                rng = np.random.default_rng(int(base_seed) + i)
                d = sp.draw_reference(rng, locs[i], augment)
                b = o.get("budget", self.budget)
                b = int(self.store.budget0[d["s"], d["y"]]) if b is None else int(np.asarray(b).reshape(-1)[i % np.size(b)])
                if o.get("sample_budget", False):  # env.py:172-177
                    if o.get("sample_budget_type", "less_than") == "less_than":
                        b = int(rng.integers(0, b + 1))
                    else:
                        b = int(rng.integers(0.5 * b, 1.5 * b + 1))
                rows.append((d["s"], d["y"], d["location_index"], d["coef_index"], sp.fips_list.index(d["info_location"])))
                budgets.append(b)
            arr = np.array(rows, dtype=np.int32).reshape(n, 5)
            s = {k: torch.from_numpy(np.ascontiguousarray(arr[:, j])).to(dev)
                 for j, k in enumerate(("loc", "year", "coef_col", "coef_idx", "list_idx"))}
            s["budget"] = torch.from_numpy(np.array(budgets, dtype=np.int32)).to(dev)
            return s
        if location is None:
            list_idx = self._randint(len(sp.fips_list), n)  # env.py:151-152
        elif isinstance(location, str):
            list_idx = torch.full((n,), sp.fips_list.index(location), dtype=torch.int64, device=dev)  # ValueError
        else:
            h = sp.draw_bulk(self._host_rng, n, location, False)["list_idx"]
            list_idx = torch.from_numpy(h.astype(np.int64)).to(dev)
        loc = self._list_to_data[list_idx]
        no_similar = None
        if augment:  # index inside the similar-county list (env.py:117; Q6)
            n_c = self._n_similar[list_idx]
            no_similar = n_c <= 0
            u = torch.rand(n, generator=self._gen, device=dev, dtype=torch.float64)
            coef_col = torch.minimum((u * n_c).long(), n_c - 1)
            # env.py:118: the drawn similar county is what info["location"] reports ...
            info_idx = self._similar_members[self._similar_start[list_idx] + coef_col.clamp(min=0)]
            if augment == "intended":  # NOT the reference: ... and the episode moves there as well
                list_idx = info_idx
                coef_col = list_idx
                loc = self._list_to_data[list_idx]
        else:
            coef_col = list_idx  # env.py:121
            info_idx = list_idx
        year = self._valid_year_idx[self._randint(len(sp.valid_years), n)]  # env.py:125
        coef_idx = self._randint(sp.n_samples, n)  # env.py:160
        if not self._all_episodes_exist or (augment and not self._augment_safe):
            bad = (loc < 0) | (year < 0)
            if augment:
                bad |= no_similar
            bad |= self._n_days_dev[(loc.clamp(min=0) * self.core.Y + year.clamp(min=0))] == 0
            if bool(bad.any()):  # the reference raises KeyError at env.py:127 / datautils.py:123
                b = int(torch.nonzero(bad)[0])
                raise KeyError((sp.fips_list[int(list_idx[b])], int(year[b])))
        list_idx = info_idx
        if self.sort_episodes and sort:
            Sw, P, Y = self.core.Sw, sp.n_samples, self.core.Y
            key = ((loc * Sw + coef_col) * P + coef_idx) * Y + year
            if list_idx is coef_col:
                # no augmentation: the key holds everything -- sort the keys alone (32-bit when they fit: half the radix
                # passes) and take the streams apart again, instead of an argsort and five gathers
                if self.core.S * Sw * P * Y < 2**31:
                    key = key.int()
                key = torch.sort(key).values.long()
                year = key % Y
                coef_idx = (key // Y) % P
                coef_col = (key // (Y * P)) % Sw
                loc = key // (Y * P * Sw)
                list_idx = coef_col
            else:
                order = torch.argsort(key)
                loc, year, coef_col, coef_idx, list_idx = (v[order] for v in (loc, year, coef_col, coef_idx, list_idx))
        s = dict(loc=loc.int(), year=year.int(), coef_col=coef_col.int(), coef_idx=coef_idx.int(),
                 list_idx=list_idx.int())
        budget = o.get("budget", self.budget)
        if budget is not None or o.get("sample_budget", False):
            if budget is None:
                b = self.core.budget0[(loc * self.core.Y + year)].long()
            elif np.ndim(budget) == 0:
                b = torch.full((n,), int(budget), dtype=torch.int64, device=dev)  # a fill kernel: graph capturable
            else:
                key = id(budget)
                if getattr(self, "_budget_cache", (None, None))[0] != key:  # uploaded once, reused by captured resets
                    self._budget_cache = (key, torch.as_tensor(
                        np.ascontiguousarray(np.broadcast_to(np.asarray(budget, dtype=np.int64), (n,)))).to(dev))
                b = self._budget_cache[1]
            if o.get("sample_budget", False):  # env.py:172-177, per env, without the sticky compounding (Q8)
                u = torch.rand(n, generator=self._gen, device=dev, dtype=torch.float64)
                if o.get("sample_budget_type", "less_than") == "less_than":
                    lo, hi = torch.zeros_like(b), b + 1
                else:
                    lo, hi = (0.5 * b.double()).long(), (1.5 * b.double() + 1).long()
                b = lo + torch.minimum((u * (hi - lo)).long(), hi - lo - 1)
            s["budget"] = b.int()
        return s

    # ---- reset --------------------------------------------------------------------------------
    def reset(self, *, seed: int | None = None, options: dict | None = None):
        """`options` keys: `location` (FIPS, or a sequence of N), `similar_climate_counties` (False, True =
        the reference's augmentation, or "intended" = the episode moves to a county drawn from the similar
        list, weather and coefficient column alike -- a documented non-parity mode),
        `budget` (int or sequence), `sample_budget`, `sample_budget_type`, `reference_rng` (True: env i
        draws from its own ``default_rng(seed + i)`` in the reference's order, like a gymnasium
        SyncVectorEnv of reference envs), or `index_streams` (dict of int arrays `loc`, `year`,
        `coef_col`, `coef_idx`[, `budget`, `list_idx`])."""
        if seed is not None:
            self._gen.manual_seed(int(seed))
            self._host_rng = np.random.default_rng(seed)
        self._reset_opts = dict(options or {})
        if self._reset_opts.get("reference_rng", False):
            if seed is None:
                raise ValueError("options['reference_rng'] needs reset(seed=...)")
            self._reset_opts["reference_seed"] = int(seed)
        self._step_no = 0
        self._reset_subset(None)
        return self._obs, self._infos()

    def _reset_subset(self, mask):
        s = self._draw(sort=mask is None)
        self._idx = s
        o = self._reset_opts
        self._slots_sorted = bool(mask is None and self.sort_episodes and "index_streams" not in o
                                  and not o.get("reference_rng", False))
        self.core.reset(self._state, self._obs, s["loc"], s["year"], s["coef_col"], s["coef_idx"],
                        budget=s.get("budget"), mask=mask, remaining=self._rem)
        if mask is None:
            self._ep_loc, self._ep_year, self._coef_col, self._list_idx = s["loc"], s["year"], s["coef_col"], s["list_idx"]
            self._atb.zero_()
        else:
            m = mask.bool()
            self._ep_loc = torch.where(m, s["loc"], self._ep_loc)
            self._ep_year = torch.where(m, s["year"], self._ep_year)
            self._coef_col = torch.where(m, s["coef_col"], self._coef_col)
            self._list_idx = torch.where(m, s["list_idx"], self._list_idx)
            self._atb &= ~m
        if self._uniform_days is not None:
            if mask is None:  # lockstep: the host knows the whole schedule
                self._term_at = None
                self._next_term = self._step_no + self._uniform_days
                return
        ends = self._step_no + self._n_days_dev[(s["loc"].long() * self.core.Y + s["year"].long())].long()
        if mask is None or self._term_at is None:
            self._term_at = ends if mask is None else torch.where(
                mask.bool(), ends, torch.full_like(ends, self._next_term))
        else:
            self._term_at = torch.where(mask.bool(), ends, self._term_at)
        self._next_term = int(self._term_at.min())  # one small sync per reset, only for ragged summers

    # ---- step -----------------------------------------------------------------------------------
    def _variant(self) -> int:
        """The step kernel for the current episodes: the caller's choice, or (AUTO) the per-step logit
        table when every env steps on its own county's coefficients and the batch is large enough to pay
        for building the table (`HeatAlertCore.choose_variant`)."""
        if self.step_variant != _lib.STEP_AUTO:
            return self.step_variant
        v = self.core.choose_variant(self.num_envs, canonical=self._canonical)
        return v | _lib.STEP_LOCAL if (v == _lib.STEP_TABLE and self._slots_sorted) else v

    def step(self, actions: torch.Tensor):
        if self._next_term is None:
            raise RuntimeError("call reset() before step()")
        if not isinstance(actions, torch.Tensor):
            actions = torch.as_tensor(np.asarray(actions), device=self.device)
        if actions.device != self.device:
            actions = actions.to(self.device)
        if actions.dtype == torch.bool:
            actions = actions.view(torch.uint8)
        if not actions.is_contiguous():
            actions = actions.contiguous()
        variant = self._variant()
        fn = self._bound.get((actions.dtype, variant))
        if fn is None:
            fn = self._bound[(actions.dtype, variant)] = self.core.bind_step(
                self._state, self._obs, self._reward, self._term.view(torch.uint8), self._rem,
                self._atb.view(torch.uint8), variant=variant, action_dtype=actions.dtype)
        fn(actions)
        if self.stats is not None:
            self.stats.update(actions, self._reward, self._term, self._atb)
        self._step_no += 1
        self._final_info = None
        if self.autoreset and self._step_no >= self._next_term:
            # SAME_STEP autoreset overwrites the info tensors of the envs it resets: keep the terminal step's
            self._final_info = {"remaining_budget": self._rem.clone(), "at_budget": self._atb.clone(),
                                "_mask": self._term.clone()}
            if self._term_at is None:
                self._reset_subset(None)
            else:
                self._reset_subset((self._term_at <= self._step_no).to(torch.uint8))
        return self._obs, self._reward, self._term, self._trunc, self._infos()

    # ---- episode-fused rollout ------------------------------------------------------------------
    def rollout(self, steps: int | None = None, actions: torch.Tensor | None = None, p_alert: float = 0.5,
                seed: int = 0, stats: torch.Tensor | None = None, policy=None, stochastic: bool = False):
        """Run up to `steps` steps (default: to the end of the summer) of every env in ONE kernel
        launch -- the `while not done: step(a)` loop of `env.py:273-276` without materialising
        observations.  `actions` is a ``uint8/bool [steps, N]`` CUDA tensor, or None for
        counter-based Bernoulli(`p_alert`) actions (`rollout_action_bits` reproduces them).
        Returns ``(episode_return float64[N], stats float64[6])``; `stats` (see
        `distributed.STAT_NAMES`) is accumulated into, so it can be carried across calls and
        all-reduced with `distributed.reduce_stats`.  Envs that finish are NOT auto-reset.

        `policy` (30 numbers: one weight per observation slot in `feature_names` order -- slot 27, the
        significance label, is ignored -- and the bias) puts the agent inside the launch instead:
        ``a = [sum_s policy[s] * obs[s] + policy[29] > 0]``, or with ``stochastic=True`` a Bernoulli draw
        with probability ``sigmoid(logit)`` from the counter-based stream of `seed`.  It needs envs at the
        start of a summer (right after `reset()`); others are left untouched."""
        if self._next_term is None:
            raise RuntimeError("call reset() before rollout()")
        if steps is None:
            steps = int(self.store.n_days.max())
        if actions is not None:
            if actions.dtype == torch.bool:
                actions = actions.view(torch.uint8)
            actions = actions.to(self.device).contiguous()
            if actions.shape != (steps, self.num_envs):
                raise ValueError(f"actions must have shape ({steps}, {self.num_envs})")
        if stats is None:
            stats = torch.zeros(15, dtype=torch.float64, device=self.device)
        ep_return = torch.zeros(self.num_envs, dtype=torch.float64, device=self.device)
        if policy is not None:
            if actions is not None:
                raise ValueError("rollout: give either `actions` or `policy`")
            theta = torch.as_tensor(policy, dtype=torch.float64).reshape(-1).to(self.device).contiguous()
            if theta.numel() != _lib.POLICY_LEN:
                raise ValueError(f"policy must hold {_lib.POLICY_LEN} numbers (29 observation slots + bias)")
            self.core.rollout_policy(self._state, steps, ep_return, theta, stochastic=stochastic, seed=seed, stats=stats)
            self._step_no += steps
            return ep_return, stats
        self.core.rollout(self._state, steps, ep_return, actions=actions, seed=seed, p_alert=p_alert, stats=stats)
        self._step_no += steps
        return ep_return, stats

    # ---- a fused MLP policy for the step-by-step loop ----------------------------------------------------
    def mlp_policy(self, w1, b1, w2, b2: float, activation: str = "relu", stochastic: bool = False, seed: int = 0):
        """`policy(obs) -> actions` as ONE kernel launch (`w2a_policy_mlp`): ``logit = w2 . act(W1 . nan_to_num(obs) + b1)
        + b2``, ``action = logit > 0`` -- or a Bernoulli(sigmoid(logit)) draw from the counter-based stream of `seed`,
        whose step counter advances with every call.  `w1` is ``[H, 29]`` (torch.nn.Linear layout), H a multiple of 8 up
        to 64; the contraction runs on the tensor cores (float16 inputs, float32 accumulation).  The returned callable has `.logits` (float32 [N], the
        last call's) and works with `graphed_rollout` (capturable; a captured stochastic policy replays the same draws
        every summer, so re-capture or use the deterministic form there)."""
        dev, n = self.device, self.num_envs
        f32 = lambda a: torch.as_tensor(a, dtype=torch.float32).to(dev).contiguous()  # noqa: E731
        w1_t, b1_t, w2_t = f32(w1), f32(b1).reshape(-1), f32(w2).reshape(-1)
        H = b1_t.numel()
        if w1_t.shape != (H, schema.OBS_DIM) or w2_t.numel() != H:
            raise ValueError(f"mlp_policy: w1 must be [{H}, {schema.OBS_DIM}], b1 and w2 [{H}]")
        act = torch.zeros(n, dtype=torch.uint8, device=dev)
        logits = torch.zeros(n, dtype=torch.float32, device=dev)
        core = self.core

        def policy(obs):
            core.policy_mlp(obs, w1_t, b1_t, w2_t, float(b2), act, logit=logits, activation=activation,
                            stochastic=stochastic, seed=seed, step=policy.calls)
            policy.calls += 1
            return act

        policy.calls = 0
        policy.logits = logits
        # what `graphed_rollout` needs to run step + policy as ONE launch per iteration (`w2a_step_mlp`); set
        # `policy.fuse_step = False` to keep the two launches
        policy.fuse_step = True
        policy.mlp = dict(w1=w1_t, b1=b1_t, w2=w2_t, b2=float(b2), activation=activation, stochastic=stochastic, seed=seed)
        policy.out = act
        return policy

    # ---- policy in the loop, replayed as one CUDA graph ------------------------------------------------
    def graphed_rollout(self, policy, steps: int | None = None, reset_at_end: bool = False):
        """Capture `steps` iterations of ``actions = policy(obs); env.step(actions)`` into ONE CUDA
        graph (the policy's torch kernels and the step kernel, no host work in between) and return
        a `GraphedRollout`; every `replay()` advances all envs by `steps` steps and returns the
        float64 return of each env over those steps.

        `policy` maps the float32 ``[N, 29]`` observation tensor to an ``[N]`` action tensor
        (bool / integer; non-zero = alert) using capturable torch ops only.  The capture must
        not run past the end of an episode; with ``reset_at_end=True`` (summers of equal length,
        `steps` = days left) the reset of all envs is captured too, so that `replay()` can be
        called forever: one replay = one summer."""
        if self._next_term is None:
            raise RuntimeError("call reset() before graphed_rollout()")
        left = self._next_term - self._step_no
        steps = left if steps is None else int(steps)
        if steps < 1 or steps > left or self._term_at is not None and steps > left:
            raise ValueError(f"graphed_rollout: steps={steps} must be in [1, {left}] (days left in the episodes)")
        if reset_at_end and (self._uniform_days is None or steps != left or "index_streams" in self._reset_opts
                             or self._reset_opts.get("reference_rng", False)):
            raise ValueError("reset_at_end needs summers of equal length, steps == days left and device-side sampling")
        augment = self._reset_opts.get("similar_climate_counties", self.similar_climate_counties)
        if reset_at_end and (not self._all_episodes_exist or (augment and not self._augment_safe)):
            raise ValueError("reset_at_end: some (county, year) episodes or similar-county lists are missing, so a "
                             "reset needs a host-side check and cannot be captured")
        return GraphedRollout(self, policy, steps, reset_at_end)

    def _infos(self):
        loc, year, lidx = self._ep_loc, self._ep_year, self._list_idx
        fips, years, flist = self.store.fips, self.store.years, self.fips_list

        def episode_index():  # env.py:156: ORIGINAL location (whose weather is played, Q7) + "_" + year
            lo, yr = loc.cpu().numpy(), year.cpu().numpy()
            return np.array([f"{fips[a]}_{years[b]}" for a, b in zip(lo, yr)], dtype=object)

        def location():  # env.py:118,120: the similar county when augmenting, else the county itself
            return np.array([flist[i] for i in lidx.cpu().numpy()], dtype=object)

        infos = {"remaining_budget": self._rem, "at_budget": self._atb,
                 "location_index": self._coef_col,       # env.py:117,121,234 (Q6: index in the similar list)
                 "location_list_index": self._list_idx,  # info["location"] as fips_list index
                 "episode_index": LazyStrings(self.num_envs, episode_index),
                 "location": LazyStrings(self.num_envs, location),
                 "feature_names": self.feature_names}
        if self._final_info is not None:
            infos["final_info"] = self._final_info
        return infos

    # ---- checkpoint / resume ----------------------------------------------------------------------
    def state_dict(self):
        """Everything a fresh env (same tables, same constructor arguments) needs to continue exactly where this
        one is: device state, observations and info tensors, the autoreset schedule, both RNGs, the options of
        the last `reset` (they drive every later autoreset) and the rollout statistics."""
        opts = {k: v for k, v in self._reset_opts.items()}
        sd = {"state": self._state.clone(), "obs": self._obs.clone(), "rem": self._rem.clone(), "atb": self._atb.clone(),
              "step_no": self._step_no, "term_at": None if self._term_at is None else self._term_at.clone(),
              "next_term": self._next_term, "canonical": self._canonical, "slots_sorted": self._slots_sorted, "reset_opts": opts,
              "ep_loc": self._ep_loc.clone(), "ep_year": self._ep_year.clone(), "coef_col": self._coef_col.clone(),
              "list_idx": self._list_idx.clone(),
              "gen": self._gen.get_state(), "host_rng": self._host_rng.bit_generator.state}
        if self.stats is not None:
            sd["stats"] = {"totals": self.stats.totals.clone(), "ep_return": self.stats.ep_return.clone()}
        return sd

    def load_state_dict(self, sd):
        self._state.copy_(sd["state"])
        self._obs.copy_(sd["obs"])
        self._rem.copy_(sd["rem"])
        self._atb.copy_(sd["atb"])
        self._step_no = int(sd["step_no"])
        self._term_at = None if sd["term_at"] is None else sd["term_at"].clone()
        self._next_term = sd["next_term"]
        self._canonical = bool(sd["canonical"])
        self._slots_sorted = bool(sd.get("slots_sorted", False))
        self._reset_opts = dict(sd["reset_opts"])
        self._ep_loc, self._ep_year = sd["ep_loc"].clone(), sd["ep_year"].clone()
        self._coef_col, self._list_idx = sd["coef_col"].clone(), sd["list_idx"].clone()
        self._gen.set_state(sd["gen"])
        self._host_rng.bit_generator.state = sd["host_rng"]
        self._final_info = None
        if self.stats is not None and "stats" in sd:
            self.stats.totals.copy_(sd["stats"]["totals"])
            self.stats.ep_return.copy_(sd["stats"]["ep_return"])

    def close(self, **kwargs):
        self.core = None
        self.closed = True


class GraphedRollout:
    """One CUDA graph of `steps` x (policy forward + step kernel); see `HeatAlertVectorEnv.graphed_rollout`."""

    def __init__(self, env: HeatAlertVectorEnv, policy, steps: int, reset_at_end: bool):
        self.env, self.steps, self.reset_at_end = env, steps, reset_at_end
        n, dev = env.num_envs, env.device
        self.returns = torch.zeros(n, dtype=torch.float64, device=dev)
        self._act = torch.zeros(n, dtype=torch.uint8, device=dev)
        step = env.core.bind_step(env._state, env._obs, env._reward, env._term.view(torch.uint8), env._rem,
                                  env._atb.view(torch.uint8), variant=env._variant(), action_dtype=torch.uint8)

        fused = getattr(policy, "fuse_step", False) and hasattr(policy, "mlp")
        if fused:
            act2 = torch.zeros(n, dtype=torch.uint8, device=dev)  # the policy's output alternates between two buffers

        def body_fused():
            # iteration k is ONE launch: the step of action k and the policy's action k + 1 from the fresh observations
            # (`w2a_step_mlp`); the observation tensor is written for whoever trains on it, never read back here
            self.returns.zero_()
            cur, nxt = policy(env._obs), act2
            for _ in range(steps):
                env.core.step_mlp(env._state, cur, env._obs, env._reward, env._term.view(torch.uint8), env._rem,
                                  env._atb.view(torch.uint8), next_action=nxt, logit=policy.logits, ep_return=self.returns,
                                  step=policy.calls, variant=env._variant(), **policy.mlp)
                policy.calls += 1
                if env.stats is not None:
                    env.stats.update(cur, env._reward, env._term, env._atb)
                cur, nxt = nxt, cur
            self._reset_tail()

        def body():
            if fused:
                return body_fused()
            self.returns.zero_()
            for _ in range(steps):
                a = policy(env._obs)
                if a.dtype == torch.uint8 and a.shape == (n,) and a.is_contiguous():
                    act = a  # already what the step kernel reads (e.g. `mlp_policy`): no conversion kernels
                else:
                    self._act.copy_(a.reshape(n).ne(0) if a.dtype != torch.bool else a.reshape(n))
                    act = self._act
                step(act)
                if env.stats is not None:
                    env.stats.update(act, env._reward, env._term, env._atb)
                self.returns.add_(env._reward)
            self._reset_tail()

        # warm up on a side stream (lazy initialisation of torch kernels and of the launch tables), on
        # a copy of the env state so that the capture starts from where the env is now
        info_t = ("_ep_loc", "_ep_year", "_coef_col", "_list_idx")
        for name in info_t:  # the captured reset copies into these: they must be tensors the env owns
            setattr(env, name, getattr(env, name).clone())
        snap = (env._state.clone(), env._obs.clone(), env._gen.get_state(), [getattr(env, k).clone() for k in info_t])
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            body()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        env._state.copy_(snap[0])
        env._obs.copy_(snap[1])
        env._gen.set_state(snap[2])
        for name, t in zip(info_t, snap[3]):
            getattr(env, name).copy_(t)
        if env.stats is not None:
            env.stats.totals.zero_()
            env.stats.ep_return.zero_()
        self.graph = torch.cuda.CUDAGraph()
        self.graph.register_generator_state(env._gen)
        with torch.cuda.graph(self.graph):
            body()
        # the capture itself does not execute anything: the env is still where it was

    def _reset_tail(self):
        """the captured reset of all envs that ends a summer (`reset_at_end`)"""
        env = self.env
        if not self.reset_at_end:
            return
        s = env._draw()
        env._idx = s
        env.core.reset(env._state, env._obs, s["loc"], s["year"], s["coef_col"], s["coef_idx"],
                       budget=s.get("budget"), remaining=env._rem)
        for dst, key in ((env._ep_loc, "loc"), (env._ep_year, "year"), (env._coef_col, "coef_col"),
                         (env._list_idx, "list_idx")):
            dst.copy_(s[key])
        env._atb.zero_()

    def replay(self) -> torch.Tensor:
        env = self.env
        if not self.reset_at_end and env._step_no + self.steps > env._next_term:
            raise RuntimeError("replay would run past the end of the episodes; reset() the env first")
        self.graph.replay()
        env._step_no += self.steps
        if self.reset_at_end:
            env._next_term = env._step_no + env._uniform_days
        return self.returns
