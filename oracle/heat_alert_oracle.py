"""CPU restatement of the reference's HeatAlertEnv reset/step path -- the parity oracle.

TEST INFRASTRUCTURE -- not part of the product.  Only `tests/`, `tests/golden/make_golden.py`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / reference legs may import this
module; nothing under `weather2alert_b200/` does.

Parity status: the reference ships no tests and no golden vectors,
so this oracle is pinned against outputs of the UNMODIFIED reference run in the build
container (`oracle/ref_harness.py`) and committed as `tests/golden/*.npz` together with
the generating script `tests/golden/make_golden.py`; `tests/test_oracle_golden.py`
replays them.

Two forms of the same algorithm, both in float64 on the *unpacked* arrays
(`feat[S,Y,T,27]`, `sig_code[S,Y,T]`, `coef[P,S,2,28]`) so that the device packing is
checked too:

* `OracleEnv` -- scalar, one env, follows `src/weather2alert/env.py` line by line
  including the numpy RNG draw order of `reset` (env.py:143-177);
* `batch_reset` / `batch_step` -- numpy-vectorised over N envs, takes the same index
  streams the CUDA reset takes.

Every function cites the reference lines it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import numpy as np

REWARD_SCALE = -1000 / 152  # env.py:221
N_FEAT = 27
OBS_DIM = 29
SLOT_LAG1, SLOT_STREAK, SLOT_REMAINING, SLOT_SIG, SLOT_2WKS = 23, 24, 25, 27, 28

# datautils.py:3-17 and :42-100 (facts, restated; "03" -> AZ is the reference's)
_WESTERN = {"AZ", "CA", "CO", "ID", "MT", "NM", "NV", "OR", "WA", "ND", "SD", "NE", "KS"}
_STATE = dict(
    (p[:2], p[2:])
    for p in (
        "01AL 02AK 03AZ 04AZ 05AR 06CA 08CO 09CT 10DE 11DC 12FL 13GA 15HI 16ID 17IL 18IN 19IA 20KS "
        "21KY 22LA 23ME 24MD 25MA 26MI 27MN 28MS 29MO 30MT 31NE 32NV 33NH 34NJ 35NM 36NY 37NC 38ND "
        "39OH 40OK 41OR 42PA 44RI 45SC 46SD 47TN 48TX 49UT 50VT 51VA 53WA 54WV 55WI 56WY 72PR 60AS "
        "66GU 69MP 78VI"
    ).split()
)


def expit(z):
    """scipy.special.expit (env.py:12) in float64."""
    return 1.0 / (1.0 + np.exp(-np.asarray(z, dtype=np.float64)))


def similar_counties(fips, conf_fips, conf_zone):
    """datautils.py:103-126: counties whose relabelled climate zone equals that of `fips`."""
    labels = []
    for f, z in zip(conf_fips, conf_zone):
        western = _STATE.get(str(f)[:2]) in _WESTERN  # datautils.py:109-110
        labels.append("Cold-West" if western else ("Cold-East" if z == "Cold" else z))  # :113-120
    if fips not in conf_fips:
        raise KeyError(fips)  # confounders.loc[fips], datautils.py:123
    zone = labels[list(conf_fips).index(fips)]
    return [f for f, lab in zip(conf_fips, labels) if lab == zone]  # datautils.py:124


def reward_f64(x27, coef_2x28, action):
    """env.py:197-221 for one env: x27 = the 27 numeric row values seen by the reward."""
    zb = float(np.dot(coef_2x28[0, :N_FEAT].astype(np.float64), x27) + float(coef_2x28[0, N_FEAT]))
    ze = float(np.dot(coef_2x28[1, :N_FEAT].astype(np.float64), x27) + float(coef_2x28[1, N_FEAT]))
    baseline = expit(zb)  # env.py:211
    effectiveness = expit(ze) * (x27[0] > 0.5)  # env.py:218
    return float(REWARD_SCALE * baseline * (1 - effectiveness * action))  # env.py:221


class OracleEnv:
    """Scalar restatement of `HeatAlertEnv` (env.py:17-262) over dense arrays.

    `fips`, `years`, `n_days`, `feat`, `sig_code`, `conf_fips`, `conf_zone` describe the
    feature store; `fips_list`, `coef` the checkpoint.  Observations are float64[29] with
    the `significance` slot as a float code (NaN = none).
    """

    def __init__(self, fips, years, n_days, feat, sig_code, conf_fips, conf_zone, fips_list, coef,
                 valid_years=None, similar_climate_counties=False, budget=None):
        self.fips = list(fips)
        self.years = [int(y) for y in years]
        self.n_days_tab, self.feat, self.sig_code = n_days, feat, sig_code
        self.conf_fips, self.conf_zone = list(conf_fips), list(conf_zone)
        self.fips_list = [str(x) for x in fips_list]  # env.py:75
        self.coef = coef
        self.n_samples = coef.shape[0]  # env.py:85
        self.valid_years = valid_years  # env.py:32
        self.similar_climate_counties = similar_climate_counties
        self.budget = budget  # env.py:34
        if self.valid_years is None:  # env.py:104-105: years in order of first appearance
            self.valid_years = list(self.years)

    def _get_episode(self, location, augment):  # env.py:107-131
        if augment:
            locations = similar_counties(location, self.conf_fips, self.conf_zone)  # :115
            locations = [x for x in locations if x in self.fips_list]  # :116
            self.location_index = int(self.rng.choice(range(len(locations))))  # :117
            self.location = locations[self.location_index]  # :118
        else:
            self.location = location
            self.location_index = self.fips_list.index(location)  # :121 (ValueError if absent)
        year = self.rng.choice(self.valid_years)  # :125
        if location not in self.fips or int(year) not in self.years:
            raise KeyError((location, year))  # merged.loc[...] :127
        s, y = self.fips.index(location), self.years.index(int(year))  # ORIGINAL location (Q7)
        if self.n_days_tab[s, y] == 0:
            raise KeyError((location, year))
        return (s, y), year

    def reset(self, location=None, similar_climate_counties=None, seed=None, budget=None,
              sample_budget=False, sample_budget_type="less_than"):  # env.py:133-184
        if seed is None:
            seed = np.random.randint(0, 10000)  # :143-144
        self.rng = np.random.default_rng(seed)  # :145
        if similar_climate_counties is None:
            similar_climate_counties = self.similar_climate_counties
        if location is None:
            location = self.rng.choice(self.fips_list)  # :151-152
        self.ep, year = self._get_episode(location, similar_climate_counties)  # :155
        self.ep_index = str(location) + "_" + str(year)  # :156
        self.n_days = int(self.n_days_tab[self.ep])  # :157
        self.coef_index = int(self.rng.integers(0, self.n_samples))  # :160
        self.attempted_alert_buffer = []
        self.actual_alert_buffer = []
        self.alert_streak = 0
        self.t = 0
        if self.budget is None:  # :167-170 (sticky afterwards, Q8)
            self.budget = int(self.feat[self.ep][0, SLOT_REMAINING]) if budget is None else budget
        if sample_budget:  # :172-177
            b = self.budget
            if sample_budget_type == "less_than":
                self.budget = int(self.rng.integers(0, b + 1))
            elif sample_budget_type == "centered":
                self.budget = int(self.rng.integers(0.5 * b, 1.5 * b + 1))
        self.remaining_budget = self.budget
        self.at_budget = False
        self.observation = self._get_obs()
        return self.observation.copy(), self._get_info()

    def _get_obs(self):  # env.py:186-195
        row = np.empty(OBS_DIM, dtype=np.float64)
        row[:N_FEAT] = self.feat[self.ep][self.t]
        row[SLOT_SIG] = self.sig_code[self.ep][self.t]
        row[SLOT_LAG1] = self.actual_alert_buffer[-1] if self.t > 0 else 0  # :190
        row[SLOT_2WKS] = sum(self.actual_alert_buffer[-14:])  # :191 (new label, Q3)
        row[SLOT_STREAK] = self.alert_streak  # :192
        row[SLOT_REMAINING] = self.budget - sum(self.actual_alert_buffer)  # :193
        return row

    def _get_reward(self, action):  # env.py:197-226
        row = self._get_obs()
        r = reward_f64(row[:N_FEAT], self.coef[self.coef_index, self.location_index], action)
        # env.py:223-224 is dead code: `action` is already the clamped one
        return r

    def _get_info(self):  # env.py:228-236
        return {
            "episode_index": self.ep_index,
            "remaining_budget": self.remaining_budget,
            "at_budget": self.at_budget,
            "location": self.location,
            "location_index": self.location_index,
        }

    def step(self, action):  # env.py:238-262
        self.attempted_alert_buffer.append(action)
        self.at_budget = sum(self.actual_alert_buffer) == self.budget  # :242
        actual_action = 0 if (action == 1 and self.at_budget) else action  # :243-246
        self.actual_alert_buffer.append(actual_action)
        if actual_action == 1:
            self.remaining_budget -= 1  # :249-250
        reward = self._get_reward(actual_action)  # :253
        done = self.t >= self.n_days - 1  # :256
        if not done:
            self.observation = self._get_obs()  # :258
            self.t += 1
            self.alert_streak = self.alert_streak + 1 if actual_action else 0  # :260
        return self.observation.copy(), reward, done, False, self._get_info()


# ---- batched form -------------------------------------------------------------------------

def batch_reset(feat, sig_code, n_days, loc, year, coef_col, coef_idx, budget=None):
    """env.py:155-184 after the RNG draws, for N envs.  `budget` None / entries < -2**30 mean
    the historical default (`ep["remaining_budget"].iloc[0]`, env.py:169)."""
    loc, year = np.asarray(loc), np.asarray(year)
    n = len(loc)
    b0 = feat[loc, year, 0, SLOT_REMAINING].astype(np.int64)
    if budget is None:
        bud = b0
    else:
        budget = np.asarray(budget, dtype=np.int64)
        bud = np.where(budget < -(2**30), b0, budget)
    st = {
        "loc": loc.astype(np.int64), "year": year.astype(np.int64),
        "col": np.asarray(coef_col, dtype=np.int64), "p": np.asarray(coef_idx, dtype=np.int64),
        "budget": bud, "t": np.zeros(n, np.int64), "streak": np.zeros(n, np.int64),
        "n_days": n_days[loc, year].astype(np.int64),
        "hist": np.zeros((n, 0), np.int64),  # actual_alert_buffer of every env
    }
    obs = np.empty((n, OBS_DIM), np.float64)
    obs[:, :N_FEAT] = feat[loc, year, 0]
    obs[:, SLOT_SIG] = sig_code[loc, year, 0]
    obs[:, SLOT_LAG1] = 0
    obs[:, SLOT_2WKS] = 0
    obs[:, SLOT_STREAK] = 0
    obs[:, SLOT_REMAINING] = bud
    st["obs"] = obs
    return st, obs.copy()


def batch_step(feat, sig_code, coef, st, action):
    """env.py:238-262 for N envs; mutates `st`; returns (obs, reward, done, remaining, at_budget)."""
    a = (np.asarray(action) != 0).astype(np.int64)
    used = st["hist"].sum(axis=1)
    at_b = used == st["budget"]  # :242
    act = np.where((a == 1) & at_b, 0, a)  # :243-246
    st["hist"] = np.concatenate([st["hist"], act[:, None]], axis=1)  # :248
    used = used + act
    t = st["t"]
    x = feat[st["loc"], st["year"], t].copy()  # row of day t (ORIGINAL location, Q7)
    x[:, SLOT_LAG1] = np.where(t > 0, act, 0)  # :190
    x[:, SLOT_STREAK] = st["streak"]  # :192
    x[:, SLOT_REMAINING] = st["budget"] - used  # :193
    w = coef[st["p"], st["col"]].astype(np.float64)  # [N,2,28]  (Q6: col may be the similar-list index)
    zb = np.einsum("nk,nk->n", w[:, 0, :N_FEAT], x) + w[:, 0, N_FEAT]
    ze = np.einsum("nk,nk->n", w[:, 1, :N_FEAT], x) + w[:, 1, N_FEAT]
    eff = expit(ze) * (x[:, 0] > 0.5)  # :218
    reward = REWARD_SCALE * expit(zb) * (1 - eff * act)  # :221
    done = t >= st["n_days"] - 1  # :256
    live = ~done
    row = np.empty((len(a), OBS_DIM), np.float64)
    row[:, :N_FEAT] = x
    row[:, SLOT_SIG] = sig_code[st["loc"], st["year"], t]
    row[:, SLOT_2WKS] = st["hist"][:, -14:].sum(axis=1)  # :191
    st["obs"][live] = row[live]  # :257-258; terminal step keeps the previous obs (Q2)
    st["t"] = np.where(live, t + 1, t)
    st["streak"] = np.where(live, np.where(act == 1, st["streak"] + 1, 0), st["streak"])  # :260
    return st["obs"].copy(), reward, done, st["budget"] - used, at_b
