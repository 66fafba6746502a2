"""Generate `tests/golden/reference_trajectories.npz` from the UNMODIFIED reference env.

Run in the build container, where /root/reference is mounted:

    python tests/golden/make_golden.py

The reference has no tests or golden vectors of its own, so parity
is pinned on these: a small synthetic fixture with the reference's schema is written as
the five files the reference constructor fetches, the reference `HeatAlertEnv`
(`/root/reference/src/weather2alert/env.py`, untouched; see `oracle/ref_harness.py` for
the two import shims) is driven through the scenarios below -- they cover SURVEY.md
appendix B and quirks Q1-Q12 -- and every observation, reward, flag and info field is
stored, together with the fixture arrays themselves so that the tests do not depend on
the generator staying bit-stable.  The GPU box has no /root/reference; tests only read
the .npz.
"""
from __future__ import annotations

import json
import os
import sys
import tempfile
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)

from oracle import ref_harness  # noqa: E402
from weather2alert_b200 import synthetic  # noqa: E402
from weather2alert_b200.store import RewardWeights  # noqa: E402

# real county codes; zones as in data/raw/DoE_climate_zones.csv of the reference
# ("BA Climate Zone").  06001 is in the data but not in the weights' county list;
# 27053 and 08031 appear in the confounders frame only.
DATA_COUNTIES = [
    ("04013", "Hot-Dry"), ("06001", "Marine"), ("06037", "Hot-Dry"), ("06073", "Hot-Dry"),
    ("12086", "Hot-Humid"), ("17031", "Cold"), ("20173", "Mixed-Humid"), ("36061", "Mixed-Humid"),
    ("39035", "Cold"), ("48201", "Hot-Humid"), ("53033", "Marine"),
]
CONF_ONLY = [("27053", "Cold"), ("08031", "Cold")]
FIPS_LIST = [f for f, _ in DATA_COUNTIES if f != "06001"]
YEARS = [2006, 2007, 2008]
P = 12


def build_fixture():
    fips = [f for f, _ in DATA_COUNTIES]
    zones = [z for _, z in DATA_COUNTIES]
    n_days = np.full((len(fips), len(YEARS)), 153, dtype=np.uint8)
    n_days[fips.index("12086"), 1] = 150  # one short summer
    store = synthetic.make_store(fips, YEARS, seed=20240611, zones=zones, n_days=n_days,
                                 p_alert=0.06, extra_conf=CONF_ONLY)
    w_std = synthetic.make_weights(FIPS_LIST, P, seed=11)
    rng = np.random.default_rng(12)
    big = w_std.coef.copy()
    # heavy-tailed coefficients like the bundled legacy checkpoint (extremes to +-50; H1)
    hit = rng.random(big.shape) < 0.08
    big[hit] = (rng.choice([-1.0, 1.0], size=hit.sum()) * rng.uniform(10, 50, size=hit.sum())).astype(np.float32)
    big[..., 17] *= 0.02
    w_big = RewardWeights(fips_list=list(FIPS_LIST), coef=big)
    return store, {"std": w_std, "large": w_big}


def scenarios():
    full = 153
    sc = []
    sc.append(dict(name="cfg1_la_seed1234", weights="std", ctor={}, episodes=[
        dict(reset=dict(location="06037", seed=1234), n_steps=full + 6, p=0.5, action_seed=7)]))
    for p in (0.1, 0.9):
        sc.append(dict(name=f"fixed_p{p}", weights="std", ctor={}, episodes=[
            dict(reset=dict(location="48201", seed=5), n_steps=full, p=p, action_seed=31),
            dict(reset=dict(location="12086", seed=6), n_steps=full, p=p, action_seed=32),
            dict(reset=dict(location="12086", seed=8), n_steps=full, p=p, action_seed=33)]))
    sc.append(dict(name="random_location", weights="std", ctor={}, episodes=[
        dict(reset=dict(seed=s), n_steps=40, p=0.3, action_seed=100 + s) for s in range(20, 32)]))
    sc.append(dict(name="augmented_fixed", weights="std", ctor={}, episodes=[
        dict(reset=dict(location=loc, similar_climate_counties=True, seed=s), n_steps=60, p=0.4, action_seed=200 + s)
        for s, loc in enumerate(["06037", "17031", "48201", "53033", "20173", "06073", "04013", "39035"])]))
    sc.append(dict(name="augmented_ctor_random", weights="std", ctor=dict(similar_climate_counties=True), episodes=[
        dict(reset=dict(seed=s), n_steps=30, p=0.5, action_seed=300 + s) for s in range(40, 52)]))
    sc.append(dict(name="ctor_budget_sticky", weights="std", ctor=dict(budget=3), episodes=[
        dict(reset=dict(location="06037", seed=1, budget=9), n_steps=30, p=0.6, action_seed=401),
        dict(reset=dict(location="17031", seed=2), n_steps=30, p=0.6, action_seed=402)]))
    sc.append(dict(name="reset_budget_first_then_ignored", weights="std", ctor={}, episodes=[
        dict(reset=dict(location="06037", seed=1, budget=5), n_steps=30, p=0.6, action_seed=411),
        dict(reset=dict(location="17031", seed=2, budget=1), n_steps=30, p=0.6, action_seed=412),
        dict(reset=dict(location="36061", seed=3), n_steps=30, p=0.6, action_seed=413)]))
    sc.append(dict(name="default_budget_sticky", weights="std", ctor={}, episodes=[
        dict(reset=dict(location="36061", seed=9), n_steps=full, p=0.2, action_seed=421),
        dict(reset=dict(location="53033", seed=10), n_steps=full, p=0.2, action_seed=422)]))
    sc.append(dict(name="sample_budget_less_than", weights="std", ctor=dict(budget=12), episodes=[
        dict(reset=dict(location="06037", seed=s, sample_budget=True), n_steps=25, p=0.7, action_seed=500 + s)
        for s in range(60, 65)]))
    sc.append(dict(name="sample_budget_centered", weights="std", ctor=dict(budget=7), episodes=[
        dict(reset=dict(seed=s, sample_budget=True, sample_budget_type="centered"), n_steps=25, p=0.7,
             action_seed=600 + s) for s in range(70, 75)]))
    sc.append(dict(name="budget_zero", weights="std", ctor=dict(budget=0), episodes=[
        dict(reset=dict(location="39035", seed=3), n_steps=20, p=0.8, action_seed=701)]))
    sc.append(dict(name="years_filter", weights="std", ctor=dict(years=[2007]), episodes=[
        dict(reset=dict(seed=s), n_steps=12, p=0.5, action_seed=800 + s) for s in range(80, 86)]))
    sc.append(dict(name="after_done_short_summer", weights="std", ctor=dict(years=[2007]), episodes=[
        dict(reset=dict(location="12086", seed=4), n_steps=150 + 8, p=0.5, action_seed=901)]))
    sc.append(dict(name="large_coefs", weights="large", ctor={}, episodes=[
        dict(reset=dict(seed=s), n_steps=full, p=0.5, action_seed=1000 + s) for s in range(90, 96)]))
    sc.append(dict(name="large_coefs_augmented", weights="large", ctor=dict(similar_climate_counties=True), episodes=[
        dict(reset=dict(seed=s), n_steps=full, p=0.3, action_seed=1100 + s) for s in range(96, 100)]))
    return sc


def main():
    warnings.simplefilter("ignore")
    store, wsets = build_fixture()
    out = {
        "fix_fips": np.array(store.fips), "fix_years": np.array(store.years),
        "fix_n_days": store.n_days, "fix_feat": store.feat, "fix_sig_code": store.sig_code,
        "fix_sig_labels": np.array(store.sig_labels),
        "fix_conf_fips": np.array(store.conf_fips), "fix_conf_zone": np.array(store.conf_zone),
        "fix_fips_list": np.array(FIPS_LIST),
    }
    dirs = {}
    for name, w in wsets.items():
        out[f"fix_coef_{name}"] = w.coef
        dirs[name] = tempfile.mkdtemp(prefix=f"w2a_golden_{name}_")
        synthetic.write_reference_files(store, w, dirs[name])

    # similar-county lists straight from the reference helper (datautils.py:103-126)
    ref_cls = ref_harness.load_reference_class()
    import weather2alert.datautils as ref_du
    import pandas as pd

    conf = pd.read_parquet(os.path.join(dirs["std"], "confounders.parquet"))
    out["similar_json"] = np.array(json.dumps({f: ref_du.get_similar_counties(f, conf) for f in store.conf_fips}))

    meta = []
    k = 0
    for sc in scenarios():
        env = ref_harness.make_reference_env(dirs[sc["weights"]], **sc["ctor"])
        sc_meta = dict(name=sc["name"], weights=sc["weights"], ctor=sc["ctor"], episodes=[],
                       valid_years=[int(y) for y in env.valid_years])
        for ep in sc["episodes"]:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                obs0, info0 = env.reset(**ep["reset"])
            actions = (np.random.default_rng(ep["action_seed"]).random(ep["n_steps"]) < ep["p"]).astype(np.int64)
            obs, rew, done, rem, atb = [], [], [], [], []
            for a in actions:
                o, r, d, trunc, info = env.step(int(a))
                assert trunc is False
                obs.append(ref_harness.encode_obs(o, store.sig_labels))
                rew.append(r); done.append(d); rem.append(int(info["remaining_budget"])); atb.append(bool(info["at_budget"]))
            key = f"ep{k:03d}"
            out[key + "_obs0"] = ref_harness.encode_obs(obs0, store.sig_labels)
            out[key + "_actions"] = actions
            out[key + "_obs"] = np.array(obs)
            out[key + "_reward"] = np.array(rew, dtype=np.float64)
            out[key + "_done"] = np.array(done)
            out[key + "_rem"] = np.array(rem)
            out[key + "_atb"] = np.array(atb)
            sc_meta["episodes"].append(dict(
                key=key, reset=ep["reset"], episode_index=str(info0["episode_index"]),
                location=str(info0["location"]), location_index=int(info0["location_index"]),
                remaining_budget0=int(info0["remaining_budget"]), at_budget0=bool(info0["at_budget"]),
                coef_index=int(env.coef_index), budget=int(env.budget), n_days=int(env.n_days),
                feature_names=list(info0["feature_names"])))
            k += 1
        meta.append(sc_meta)
        print(sc["name"], len(sc["episodes"]), "episodes")
    out["meta_json"] = np.array(json.dumps(meta))
    path = os.path.join(HERE, "reference_trajectories.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB;", k, "episodes")


if __name__ == "__main__":
    main()
