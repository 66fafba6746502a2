"""The three oracle forms (scalar Python, batched numpy, C/OpenMP) against each other on random
cases beyond the golden file -- hypothesis picks shapes, budgets, summer lengths and action rates."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import c_oracle
from oracle import heat_alert_oracle as orc
from weather2alert_b200 import synthetic


@settings(max_examples=12, deadline=None)
@given(seed=st.integers(0, 10_000), n_counties=st.integers(1, 5), n_years=st.integers(1, 3),
       p_samples=st.integers(1, 4), p_act=st.sampled_from([0.0, 0.2, 0.6, 1.0]), steps=st.integers(1, 40))
def test_oracles_agree(seed, n_counties, n_years, p_samples, p_act, steps):
    rng = np.random.default_rng(seed)
    fips, zones = synthetic.synthetic_counties(n_counties, seed=seed, must_include=())
    n_days = rng.integers(1, 30, (n_counties, n_years)).astype(np.uint8)
    store = synthetic.make_store(fips, list(range(2006, 2006 + n_years)), seed=seed + 1, zones=zones, n_days=n_days,
                                 p_alert=0.2)
    w = synthetic.make_weights(fips, p_samples, seed=seed + 2, scale=1.0)
    n = 9
    loc, year = rng.integers(0, n_counties, n), rng.integers(0, n_years, n)
    col, p = rng.integers(0, n_counties, n), rng.integers(0, p_samples, n)
    budget = rng.integers(-1, 6, n)
    acts = (rng.random((steps, n)) < p_act).astype(np.int64)

    bst, bobs = orc.batch_reset(store.feat, store.sig_code, store.n_days, loc, year, col, p, budget)
    use_c = c_oracle.available()
    if use_c:
        co = c_oracle.COracle(store.feat, store.sig_code, store.n_days, w.coef, n)
        np.testing.assert_array_equal(co.reset(loc, year, col, p, budget), bobs)
    # scalar envs: feed the drawn indices directly (the RNG draw order is covered by the golden tests)
    envs = []
    for i in range(n):
        e = orc.OracleEnv(store.fips, store.years, store.n_days, store.feat, store.sig_code, store.conf_fips,
                          store.conf_zone, w.fips_list, w.coef, budget=int(budget[i]))
        e.rng = np.random.default_rng(0)
        e.ep, e.n_days = (int(loc[i]), int(year[i])), int(store.n_days[loc[i], year[i]])
        e.location_index, e.coef_index = int(col[i]), int(p[i])
        e.attempted_alert_buffer, e.actual_alert_buffer, e.alert_streak, e.t = [], [], 0, 0
        e.remaining_budget, e.at_budget, e.ep_index, e.location = e.budget, False, "x", "x"
        e.observation = e._get_obs()
        np.testing.assert_array_equal(e.observation, bobs[i])
        envs.append(e)
    for k in range(steps):
        bo, br, bd, brem, batb = orc.batch_step(store.feat, store.sig_code, w.coef, bst, acts[k])
        if use_c:
            cobs, crew, cdone, crem, catb = co.step(acts[k])
            np.testing.assert_array_equal(cobs, bo)
            np.testing.assert_allclose(crew, br, rtol=1e-12, atol=1e-300)
            np.testing.assert_array_equal(cdone.astype(bool), bd)
            np.testing.assert_array_equal(crem, brem)
            np.testing.assert_array_equal(catb.astype(bool), batb)
        for i, e in enumerate(envs):
            o, r, d, _, info = e.step(int(acts[k, i]))
            np.testing.assert_array_equal(o, bo[i])
            assert r == pytest.approx(br[i], rel=1e-12, abs=1e-300) and d == bool(bd[i])
            assert info["remaining_budget"] == brem[i] and info["at_budget"] == bool(batb[i])
