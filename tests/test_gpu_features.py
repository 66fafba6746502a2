"""Device feature engineering (`w2a_engineer_features`) against the reference's own outputs
(tests/golden/feature_engineering.npz) and against the oracle on a bigger random frame."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "feature_engineering.npz")
# feature column -> golden key, in schema.REWARD_FEATURES order (bspline columns are checked separately)
EXACT = ("heat_qi", "heat_qi_above_25", "heat_qi_above_75", "hi_max", "hi_max_above_25", "hi_max_above_75",
         "hi_max*heat_qi", "hi_max_above_25*heat_qi", "hi_max_above_75*heat_qi", "weekend", "holiday", "dos", "alert",
         "alerts_2wks", "alert_lag1", "alert_streak", "remaining_budget", "issued_in_advance")
ROLLING = ("heat_qi_3d", "excess_heat_3d", "excess_heat_3d*heat_qi", "heat_qi_7d", "excess_heat_7d", "excess_heat_7d*heat_qi")


def _starts(fips):
    return np.r_[0, np.flatnonzero(fips[1:] != fips[:-1]) + 1, len(fips)].astype(np.int64)


def test_device_features_match_the_reference_statements():
    from oracle import feature_oracle as fo
    from weather2alert_b200 import schema
    from weather2alert_b200.features import FeatureEngineer

    z = np.load(GOLDEN)
    start = _starts(z["in_fips"])
    feat, year = FeatureEngineer("cuda:0").engineer(start, z["in_date_days"], z["in_himax_c"], z["in_alert"],
                                                    z["in_issued_in_advance"], z["in_holiday"])
    feat, year = feat.cpu().numpy(), year.cpu().numpy()
    np.testing.assert_array_equal(year, z["out_year"])
    col = {name: i for i, name in enumerate(schema.REWARD_FEATURES)}
    for name in EXACT:  # integer columns and float64 columns without accumulation: bit-exact
        want = z["in_alert"] if name == "alert" else z["out_" + name]
        np.testing.assert_array_equal(feat[:, col[name]], want.astype(np.float64), err_msg=name)
    for name in ROLLING:  # pandas' online Kahan window sums vs direct ones
        np.testing.assert_allclose(feat[:, col[name]], z["out_" + name], rtol=1e-12, atol=1e-12, err_msg=name)
    # B-splines (:198-210; patsy absent -> checked against the oracle's restatement, itself checked against scipy)
    cols, _, _ = fo.bspline_dos(z["out_dos"])
    np.testing.assert_allclose(feat[:, 18:21], cols, rtol=1e-9, atol=1e-11)


@pytest.mark.parametrize("seed,S,Y", [(1, 40, 14), (2, 3, 1)])
def test_device_features_match_the_oracle_on_random_frames(seed, S, Y):
    """Bigger frames than the golden: 14 summers per county (2 142 rows, the real maximum), NaN heat index values."""
    from oracle import feature_oracle as fo
    from weather2alert_b200 import schema
    from weather2alert_b200.features import FeatureEngineer, holiday_flags

    rng = np.random.default_rng(seed)
    days, sizes = [], []
    for c in range(S):
        d = []
        for y in range(2006, 2006 + Y):
            first = (np.datetime64(f"{y}-05-01") - np.datetime64("1970-01-01")).astype(int)
            nd = 153 if rng.random() < 0.8 else int(rng.integers(1, 153))
            d.append(first + np.arange(nd))
        d = np.concatenate(d)
        days.append(d)
        sizes.append(len(d))
    days = np.concatenate(days).astype(np.int32)
    n = len(days)
    start = np.r_[0, np.cumsum(sizes)].astype(np.int64)
    himax = np.round(rng.normal(28, 5, n), 1)
    himax[rng.random(n) < 0.002] = np.nan
    alert = rng.random(n) < 0.06
    issued = np.where(alert, rng.integers(0, 6, n) / 2.0, np.nan)
    hol = holiday_flags(days)
    feat, year = FeatureEngineer("cuda:0").engineer(start, days, himax, alert, issued, hol)
    feat, year = feat.cpu().numpy(), year.cpu().numpy()
    ref = fo.engineer(start, days, himax, alert, issued)
    col = {name: i for i, name in enumerate(schema.REWARD_FEATURES)}
    np.testing.assert_array_equal(year, ref["year"])
    for name in EXACT:
        if name in ("holiday", "alert"):
            continue
        np.testing.assert_array_equal(feat[:, col[name]], ref[name].astype(np.float64), err_msg=name)
    for name in ROLLING:
        np.testing.assert_allclose(feat[:, col[name]], ref[name], rtol=1e-12, atol=1e-12, err_msg=name)
    np.testing.assert_array_equal(feat[:, col["holiday"]], hol)
    cols, _, _ = fo.bspline_dos(ref["dos"])
    np.testing.assert_allclose(feat[:, 18:21], cols, rtol=1e-9, atol=1e-11)


def test_store_from_raw_round_trip_into_the_env():
    """Raw rows -> device features -> FeatureStore -> HeatAlertVectorEnv: the env's first observations are the
    engineered day-0 rows, and the default budget is the summer's alert count minus day 0 (env.py:169, Q9)."""
    from weather2alert_b200 import HeatAlertVectorEnv, schema, synthetic
    from weather2alert_b200.features import store_from_raw

    z = np.load(GOLDEN)
    sig = np.where(z["in_alert"], "W", None)
    store = store_from_raw(z["in_fips"], z["in_date_days"], z["in_himax_c"], z["in_alert"], z["in_issued_in_advance"],
                           significance=sig, holiday=z["in_holiday"], conf_fips=sorted(set(z["in_fips"])),
                           conf_zone=["Hot-Humid"] * 7)
    assert store.fips == sorted(set(z["in_fips"])) and store.years == [2006, 2007, 2008, 2009]
    assert store.n_days[1, 1] == 0 and store.n_days[2, 2] == 120 and store.n_days[3, 0] == 143 and store.n_days[0, 0] == 153
    rows = (z["in_fips"] == store.fips[0]) & (z["out_year"] == 2006)
    np.testing.assert_array_equal(store.feat[0, 0, :153, schema.DATA_COLUMNS.index("heat_qi")], z["out_heat_qi"][rows])
    assert store.budget0[0, 0] == z["out_remaining_budget"][rows][0]
    weights = synthetic.make_weights(store.fips, 4, seed=1)
    env = HeatAlertVectorEnv(64, weights=weights, store=store, seed=3, years=[2006, 2008, 2009])
    obs, info = env.reset()
    lo, yr = env._ep_loc.cpu().numpy(), env._ep_year.cpu().numpy()
    np.testing.assert_array_equal(obs[:, :23].cpu().numpy(), store.feat[lo, yr, 0, :23].astype(np.float32))
    np.testing.assert_array_equal(info["remaining_budget"].cpu().numpy(), store.budget0[lo, yr])
    obs, rew, term, _, _ = env.step(torch.zeros(64, dtype=torch.uint8, device="cuda"))
    assert bool(torch.isfinite(rew).all())
