"""The oracle (oracle/heat_alert_oracle.py) against the outputs of the unmodified reference."""
import numpy as np
import pytest

from oracle import heat_alert_oracle as orc

REWARD_RTOL = 1e-12  # float64 vs float64: only the summation order differs


def _oracle_env(g, sc):
    ctor = sc["ctor"]
    return orc.OracleEnv(g.fips, g.years, g.n_days, g.feat, g.sig_code, g.conf_fips, g.conf_zone,
                         g.fips_list, g.coef[sc["weights"]], valid_years=ctor.get("years"),
                         similar_climate_counties=ctor.get("similar_climate_counties", False),
                         budget=ctor.get("budget"))


def test_scalar_oracle_replays_every_reference_scenario(golden):
    n_steps = 0
    for sc in golden.meta:
        env = _oracle_env(golden, sc)
        assert [int(y) for y in env.valid_years] == sc["valid_years"]
        for ep in sc["episodes"]:
            ref = golden.episode(ep["key"])
            obs, info = env.reset(**ep["reset"])
            np.testing.assert_array_equal(obs, ref["obs0"], err_msg=f"{sc['name']} {ep['key']} reset obs")
            assert info["episode_index"] == ep["episode_index"]
            assert str(info["location"]) == ep["location"]
            assert info["location_index"] == ep["location_index"]
            assert info["remaining_budget"] == ep["remaining_budget0"]
            assert info["at_budget"] == ep["at_budget0"]
            assert env.coef_index == ep["coef_index"] and env.budget == ep["budget"] and env.n_days == ep["n_days"]
            for i, a in enumerate(ref["actions"]):
                obs, r, done, trunc, info = env.step(int(a))
                where = f"{sc['name']} {ep['key']} step {i}"
                np.testing.assert_array_equal(obs, ref["obs"][i], err_msg=where)
                assert r == pytest.approx(ref["reward"][i], rel=REWARD_RTOL, abs=1e-300), where
                assert done == bool(ref["done"][i]) and trunc is False, where
                assert info["remaining_budget"] == ref["rem"][i] and info["at_budget"] == bool(ref["atb"][i]), where
                n_steps += 1
    assert n_steps > 4000


def test_batched_oracle_matches_reference(golden):
    """All 74 reference episodes as one batch through batch_reset/batch_step."""
    for wname in ("std", "large"):
        eps = [(sc, ep) for sc, ep in golden.all_episodes() if sc["weights"] == wname]
        loc = [golden.fips.index(ep["episode_index"].split("_")[0]) for _, ep in eps]
        year = [golden.years.index(int(ep["episode_index"].split("_")[1])) for _, ep in eps]
        col = [ep["location_index"] for _, ep in eps]
        p = [ep["coef_index"] for _, ep in eps]
        bud = [ep["budget"] for _, ep in eps]
        refs = [golden.episode(ep["key"]) for _, ep in eps]
        st, obs = orc.batch_reset(golden.feat, golden.sig_code, golden.n_days, loc, year, col, p, bud)
        np.testing.assert_array_equal(obs, np.stack([r["obs0"] for r in refs]))
        horizon = max(len(r["actions"]) for r in refs)
        for i in range(horizon):
            act = np.array([r["actions"][i] if i < len(r["actions"]) else 0 for r in refs])
            obs, rew, done, rem, atb = orc.batch_step(golden.feat, golden.sig_code, golden.coef[wname], st, act)
            for j, r in enumerate(refs):
                if i >= len(r["actions"]):
                    continue
                np.testing.assert_array_equal(obs[j], r["obs"][i], err_msg=f"{eps[j][1]['key']} step {i}")
                assert rew[j] == pytest.approx(r["reward"][i], rel=REWARD_RTOL, abs=1e-300)
                assert bool(done[j]) == bool(r["done"][i])
                assert int(rem[j]) == int(r["rem"][i]) and bool(atb[j]) == bool(r["atb"][i])


def test_similar_counties_match_reference_helper(golden):
    for f, want in golden.similar.items():
        assert orc.similar_counties(f, golden.conf_fips, golden.conf_zone) == want
    with pytest.raises(KeyError):
        orc.similar_counties("99999", golden.conf_fips, golden.conf_zone)


def test_c_oracle_matches_reference(golden):
    """oracle/c/heat_oracle.c (the cpu_baseline port) replays the same 74 reference episodes."""
    from oracle import c_oracle

    if not c_oracle.available():
        pytest.skip("oracle/_build/libheat_oracle.so not built (run __graft_entry__.build())")
    for wname in ("std", "large"):
        eps = [ep for sc, ep in golden.all_episodes() if sc["weights"] == wname]
        refs = [golden.episode(ep["key"]) for ep in eps]
        co = c_oracle.COracle(golden.feat, golden.sig_code, golden.n_days, golden.coef[wname], len(eps))
        obs = co.reset([golden.fips.index(ep["episode_index"].split("_")[0]) for ep in eps],
                       [golden.years.index(int(ep["episode_index"].split("_")[1])) for ep in eps],
                       [ep["location_index"] for ep in eps], [ep["coef_index"] for ep in eps],
                       [ep["budget"] for ep in eps])
        np.testing.assert_array_equal(obs, np.stack([r["obs0"] for r in refs]))
        for i in range(max(len(r["actions"]) for r in refs)):
            act = np.array([r["actions"][i] if i < len(r["actions"]) else 0 for r in refs])
            obs, rew, done, rem, atb = co.step(act)
            for j, r in enumerate(refs):
                if i < len(r["actions"]):
                    np.testing.assert_array_equal(obs[j], r["obs"][i])
                    assert rew[j] == pytest.approx(r["reward"][i], rel=REWARD_RTOL, abs=1e-300)
                    assert bool(done[j]) == bool(r["done"][i]) and rem[j] == r["rem"][i] and bool(atb[j]) == bool(r["atb"][i])
