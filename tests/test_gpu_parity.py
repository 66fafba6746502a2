"""Parity of the CUDA path (through the C ABI) with the reference's golden trajectories and
with the oracle.  Bar (BASELINE.json north_star): observations, counters and flags
bit-exact; rewards within 1e-5 relative in fp32."""
import numpy as np
import pytest

from conftest import assert_obs_equal

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

REWARD_RTOL = 1e-5   # north_star tolerance
REWARD_ATOL = 1e-10  # rewards that underflow towards 0 (sigmoid saturated) have no relative meaning
VARIANTS = {"thread": 1, "coop": 2, "table": 3}


@pytest.fixture(scope="module")
def cores(golden):
    from weather2alert_b200.core import HeatAlertCore

    store = golden.store()
    return {name: HeatAlertCore(store, golden.weights(name), "cuda:0") for name in ("std", "large")}


def _dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def _golden_batch(golden, wname):
    eps = [ep for sc, ep in golden.all_episodes() if sc["weights"] == wname]
    idx = dict(
        loc=np.array([golden.fips.index(ep["episode_index"].split("_")[0]) for ep in eps], np.int32),
        year=np.array([golden.years.index(int(ep["episode_index"].split("_")[1])) for ep in eps], np.int32),
        coef_col=np.array([ep["location_index"] for ep in eps], np.int32),
        coef_idx=np.array([ep["coef_index"] for ep in eps], np.int32),
        budget=np.array([ep["budget"] for ep in eps], np.int32),
    )
    return eps, idx, [golden.episode(ep["key"]) for ep in eps]


@pytest.mark.parametrize("variant", list(VARIANTS))
@pytest.mark.parametrize("wname", ["std", "large"])
def test_reference_trajectories_as_one_batch(golden, cores, wname, variant):
    """Every episode the unmodified reference produced, replayed in one batch per step."""
    core = cores[wname]
    eps, idx, refs = _golden_batch(golden, wname)
    n = len(eps)
    state = core.alloc_state(n)
    obs = torch.full((n, 29), -7.0, device="cuda")
    rem = torch.zeros(n, dtype=torch.int32, device="cuda")
    core.reset(state, obs, *(_dev(idx[k]) for k in ("loc", "year", "coef_col", "coef_idx")),
               budget=_dev(idx["budget"]), remaining=rem)
    assert_obs_equal(obs.cpu().numpy(), np.stack([r["obs0"] for r in refs]), "reset obs")
    np.testing.assert_array_equal(rem.cpu().numpy(), idx["budget"])
    reward = torch.zeros(n, device="cuda")
    term = torch.zeros(n, dtype=torch.uint8, device="cuda")
    atb = torch.zeros(n, dtype=torch.uint8, device="cuda")
    horizon = max(len(r["actions"]) for r in refs)
    worst = 0.0
    for i in range(horizon):
        live = np.array([i < len(r["actions"]) for r in refs])
        act = np.array([r["actions"][i] if i < len(r["actions"]) else 0 for r in refs], np.uint8)
        core.step(state, _dev(act), obs, reward, term, rem, atb, variant=VARIANTS[variant])
        o, rw, tm, rm, ab = (x.cpu().numpy() for x in (obs, reward, term, rem, atb))
        want_obs = np.stack([r["obs"][min(i, len(r["actions"]) - 1)] for r in refs])
        assert_obs_equal(o[live], want_obs[live], f"obs step {i}")
        want_r = np.array([r["reward"][i] if i < len(r["actions"]) else 0.0 for r in refs])
        np.testing.assert_allclose(rw[live], want_r[live], rtol=REWARD_RTOL, atol=REWARD_ATOL, err_msg=f"step {i}")
        worst = max(worst, float(np.max(np.abs(rw[live] - want_r[live]) / np.maximum(np.abs(want_r[live]), 1e-30))))
        np.testing.assert_array_equal(tm[live].astype(bool), np.array([r["done"][i] for r in refs if i < len(r["actions"])]))
        np.testing.assert_array_equal(rm[live], np.array([r["rem"][i] for r in refs if i < len(r["actions"])]))
        np.testing.assert_array_equal(ab[live].astype(bool), np.array([r["atb"][i] for r in refs if i < len(r["actions"])]))
    assert worst < REWARD_RTOL


def test_single_env_facade_replays_every_reference_scenario(golden):
    """`HeatAlertEnv` with the reference's own call sequence: ctor kwargs, reset kwargs
    (seeded RNG draws, sticky budgets, budget sampling), step, info dict."""
    from weather2alert_b200 import HeatAlertEnv

    store = golden.store()
    for sc in golden.meta:
        env = HeatAlertEnv(weights=golden.weights(sc["weights"]), store=store, device="cuda:0", **sc["ctor"])
        for ep in sc["episodes"]:
            ref = golden.episode(ep["key"])
            obs, info = env.reset(**ep["reset"])
            where = f"{sc['name']} {ep['key']}"
            assert obs.dtype == np.float32 and obs.shape == (29,)
            assert_obs_equal(obs, ref["obs0"], where)
            assert info["episode_index"] == ep["episode_index"] and str(info["location"]) == ep["location"]
            assert info["location_index"] == ep["location_index"]
            assert info["remaining_budget"] == ep["remaining_budget0"] and info["at_budget"] is False
            assert info["feature_names"] == ep["feature_names"]
            assert env.coef_index == ep["coef_index"] and env.budget == ep["budget"] and env.n_days == ep["n_days"]
            for i, a in enumerate(ref["actions"]):
                obs, r, done, trunc, info = env.step(int(a))
                assert_obs_equal(obs, ref["obs"][i], f"{where} step {i}")
                assert isinstance(r, float) and r == pytest.approx(ref["reward"][i], rel=REWARD_RTOL, abs=REWARD_ATOL)
                assert done is bool(ref["done"][i]) and trunc is False
                assert info["remaining_budget"] == ref["rem"][i] and info["at_budget"] is bool(ref["atb"][i])
            assert env.t == min(len(ref["actions"]), ep["n_days"] - 1)


def test_single_env_errors(golden):
    from weather2alert_b200 import HeatAlertEnv

    env = HeatAlertEnv(weights=golden.weights("std"), store=golden.store(), seed=1234)
    obs, info = env.reset(location="06037")  # ctor seed used (README form)
    assert info["episode_index"] == golden.meta[0]["episodes"][0]["episode_index"]
    with pytest.raises(ValueError):
        env.reset(location="99999")
    with pytest.raises(KeyError):
        HeatAlertEnv(weights=golden.weights("std"), store=golden.store(), years=[1999]).reset(location="06037")
    assert env.action_space.n == 2 and env.observation_space.shape == (29,)


def _random_case(seed, n, n_counties=40, n_years=3, p_samples=20, augment=False, short=True):
    from weather2alert_b200 import synthetic

    rng = np.random.default_rng(seed)
    fips, zones = synthetic.synthetic_counties(n_counties, seed=seed)
    n_days = np.full((n_counties, n_years), 153, np.uint8)
    if short:
        n_days[rng.integers(0, n_counties, 5), rng.integers(0, n_years, 5)] = rng.integers(1, 150, 5)
    store = synthetic.make_store(fips, list(range(2006, 2006 + n_years)), seed=seed + 1, zones=zones, n_days=n_days,
                                 p_alert=0.08)
    weights = synthetic.make_weights(fips, p_samples, seed=seed + 2)
    idx = dict(loc=rng.integers(0, n_counties, n), year=rng.integers(0, n_years, n),
               coef_col=rng.integers(0, n_counties, n) if augment else None,
               coef_idx=rng.integers(0, p_samples, n))
    if idx["coef_col"] is None:
        idx["coef_col"] = idx["loc"].copy()
    budget = rng.integers(-1, 12, n)
    budget[rng.random(n) < 0.3] = -(2**31)  # historical default
    idx["budget"] = budget
    return store, weights, {k: v.astype(np.int32) for k, v in idx.items()}


@pytest.mark.parametrize("variant", list(VARIANTS))
@pytest.mark.parametrize("n,augment", [(1, False), (33, True), (4096, False), (20011, True)])
def test_random_batches_against_the_oracle(n, augment, variant):
    """Seeded random batches (ragged sizes, short summers, budgets incl. 0/-1/default,
    steps past the end) against oracle/heat_alert_oracle.batch_step."""
    from oracle import heat_alert_oracle as orc
    from weather2alert_b200.core import HeatAlertCore

    store, weights, idx = _random_case(1000 + n, n, augment=augment)
    core = HeatAlertCore(store, weights, "cuda:0")
    state = core.alloc_state(n)
    obs = torch.zeros((n, 29), device="cuda")
    rem = torch.zeros(n, dtype=torch.int32, device="cuda")
    core.reset(state, obs, *(_dev(idx[k]) for k in ("loc", "year", "coef_col", "coef_idx")),
               budget=_dev(idx["budget"]), remaining=rem)
    st, o_ref = orc.batch_reset(store.feat, store.sig_code, store.n_days, idx["loc"], idx["year"], idx["coef_col"],
                                idx["coef_idx"], idx["budget"])
    assert_obs_equal(obs.cpu().numpy(), o_ref)
    np.testing.assert_array_equal(rem.cpu().numpy(), st["budget"])
    reward = torch.zeros(n, device="cuda")
    term = torch.zeros(n, dtype=torch.uint8, device="cuda")
    atb = torch.zeros(n, dtype=torch.uint8, device="cuda")
    rng = np.random.default_rng(5)
    steps = 170 if n <= 4096 else 60
    for i in range(steps):
        p = (0.05, 0.5, 0.95)[i % 3]
        act = (rng.random(n) < p).astype(np.int64)
        core.step(state, _dev(act), obs, reward, term, rem, atb, variant=VARIANTS[variant])
        o_ref, r_ref, d_ref, rem_ref, atb_ref = orc.batch_step(store.feat, store.sig_code, weights.coef, st, act)
        assert_obs_equal(obs.cpu().numpy(), o_ref, f"step {i}")
        np.testing.assert_allclose(reward.cpu().numpy(), r_ref, rtol=REWARD_RTOL, atol=REWARD_ATOL)
        np.testing.assert_array_equal(term.cpu().numpy().astype(bool), d_ref)
        np.testing.assert_array_equal(rem.cpu().numpy(), rem_ref)
        np.testing.assert_array_equal(atb.cpu().numpy().astype(bool), atb_ref)


def test_action_dtypes_masked_reset_and_null_outputs():
    from weather2alert_b200.core import HeatAlertCore

    n = 777
    store, weights, idx = _random_case(77, n)
    core = HeatAlertCore(store, weights, "cuda:0")
    d = {k: _dev(v) for k, v in idx.items()}
    act = (np.random.default_rng(0).random((6, n)) < 0.5)
    outs = {}
    for dt in (torch.uint8, torch.int32, torch.int64, torch.bool):
        state = core.alloc_state(n)
        obs = torch.zeros((n, 29), device="cuda")
        core.reset(state, obs, d["loc"], d["year"], d["coef_col"], d["coef_idx"], budget=d["budget"])
        reward = torch.zeros(n, device="cuda")
        term = torch.zeros(n, dtype=torch.uint8, device="cuda")
        for i in range(6):
            core.step(state, torch.from_numpy(act[i]).cuda().to(dt), obs, reward, term)  # remaining/at_budget NULL
        outs[dt] = (obs.clone(), reward.clone(), state.clone())
    for dt in (torch.int32, torch.int64, torch.bool):
        for a, b in zip(outs[torch.uint8], outs[dt]):
            assert torch.equal(a.view(torch.uint8), b.view(torch.uint8))
    # masked reset touches only the selected envs
    obs0, _, state0 = outs[torch.uint8]
    state, obs = state0.clone(), obs0.clone()
    mask = torch.zeros(n, dtype=torch.uint8, device="cuda")
    mask[::3] = 1
    core.reset(state, obs, d["loc"], d["year"], d["coef_col"], d["coef_idx"], budget=d["budget"], mask=mask)
    keep = (mask == 0)
    assert torch.equal(obs[keep].view(torch.int32), obs0[keep].view(torch.int32))
    fresh_state = core.alloc_state(n)
    fresh_obs = torch.zeros((n, 29), device="cuda")
    core.reset(fresh_state, fresh_obs, d["loc"], d["year"], d["coef_col"], d["coef_idx"], budget=d["budget"])
    sel = mask.bool()
    assert torch.equal(obs[sel].view(torch.int32), fresh_obs[sel].view(torch.int32))
    s_new, s_old, s_fresh = (x.view(2, n, 8) for x in (state, state0, fresh_state))
    assert torch.equal(s_new[:, keep], s_old[:, keep]) and torch.equal(s_new[:, sel], s_fresh[:, sel])


@pytest.mark.parametrize("augment", [False, True])
def test_host_buffer_entry_points_equal_device_path(augment):
    """w2a_reset_host / w2a_step_host against the device entry points, bit for bit.  The library picks the step variant
    from the host index streams: the tabulated logits when every env is on its own county's column, else COOP."""
    from weather2alert_b200 import _lib
    from weather2alert_b200.core import HeatAlertCore

    n = 150_000  # three pipelined chunks
    store, weights, idx = _random_case(99, n, short=False, augment=augment)
    core = HeatAlertCore(store, weights, "cuda:0")
    canonical = bool((idx["coef_col"] == idx["loc"]).all())  # _random_case: the weights list the store's counties in order
    assert canonical == (not augment)
    variant = _lib.STEP_TABLE if canonical else _lib.STEP_COOP
    core.host_envs_create(n)
    obs_h = np.zeros((n, 29), np.float32)
    rem_h = np.zeros(n, np.int32)
    core.reset_host(idx["loc"], idx["year"], idx["coef_col"], idx["coef_idx"], obs_h, budget=idx["budget"], remaining=rem_h)
    state = core.alloc_state(n)
    obs = torch.zeros((n, 29), device="cuda")
    d = {k: _dev(v) for k, v in idx.items()}
    core.reset(state, obs, d["loc"], d["year"], d["coef_col"], d["coef_idx"], budget=d["budget"])
    np.testing.assert_array_equal(obs_h.view(np.int32), obs.cpu().numpy().view(np.int32))
    reward = torch.zeros(n, device="cuda")
    term = torch.zeros(n, dtype=torch.uint8, device="cuda")
    rem = torch.zeros(n, dtype=torch.int32, device="cuda")
    atb = torch.zeros(n, dtype=torch.uint8, device="cuda")
    rew_h, term_h, atb_h = np.zeros(n, np.float32), np.zeros(n, np.uint8), np.zeros(n, np.uint8)
    rng = np.random.default_rng(3)
    for _ in range(5):
        act = (rng.random(n) < 0.5).astype(np.uint8)
        core.step_host(act, obs_h, rew_h, term_h, rem_h, atb_h)
        core.step(state, _dev(act), obs, reward, term, rem, atb, variant=variant)
        np.testing.assert_array_equal(obs_h.view(np.int32), obs.cpu().numpy().view(np.int32))
        np.testing.assert_array_equal(rew_h.view(np.int32), reward.cpu().numpy().view(np.int32))
        np.testing.assert_array_equal(term_h, term.cpu().numpy())
        np.testing.assert_array_equal(rem_h, rem.cpu().numpy())
        np.testing.assert_array_equal(atb_h, atb.cpu().numpy())


def test_abi_error_paths():
    import ctypes as C

    from weather2alert_b200 import _lib

    lib = _lib.lib()
    ctx = C.c_void_p()
    assert lib.w2a_create(9999, C.byref(ctx)) == -1 and b"not in" in lib.w2a_last_error(None)
    assert lib.w2a_create(0, C.byref(ctx)) == 0
    assert lib.w2a_step(ctx, None, 8, None, 0, None, None, None, None, None, None, 0) == -3  # tables missing
    assert b"w2a_set_features" in lib.w2a_last_error(ctx)
    assert lib.w2a_step_host(ctx, None, None, None, None, None, None) == -3
    assert lib.w2a_set_coefs(ctx, None, 1, 1) == -1
    nul = [None] * 6
    assert lib.w2a_step_mlp(ctx, None, 8, None, *nul, 3, None, None, None, 0.0, 8, 0, 0, 0, 0, None, None, None) == -3  # tables missing
    assert lib.w2a_policy_mlp(ctx, None, 8, None, None, None, None, 0.0, 8, 0, 0, 0, 0, None, None) == -1  # NULL arguments
    assert lib.w2a_destroy(ctx) == 0


@pytest.mark.parametrize("S,Y,P,T", [(1, 1, 1, 153), (3, 2, 5, 7), (2, 1, 2, 255)])
def test_degenerate_table_shapes(S, Y, P, T):
    """One county / one year / one posterior sample, very short and maximum-length summers."""
    from oracle import heat_alert_oracle as orc
    from weather2alert_b200 import schema, synthetic
    from weather2alert_b200.core import HeatAlertCore
    from weather2alert_b200.store import FeatureStore

    rng = np.random.default_rng(S * 100 + T)
    fips, zones = synthetic.synthetic_counties(S, seed=1)
    base = synthetic.make_store(fips, list(range(2006, 2006 + Y)), seed=2, zones=zones)
    if T != schema.MAX_DAYS:  # rebuild with a different day count
        reps = -(-T // schema.MAX_DAYS)
        feat = np.tile(base.feat, (1, 1, reps, 1))[:, :, :T].copy()
        sig = np.tile(base.sig_code, (1, 1, reps))[:, :, :T].copy()
        store = FeatureStore(fips=base.fips, years=base.years, n_days=np.full((S, Y), T, np.uint8), feat=feat,
                             sig_code=sig, sig_labels=base.sig_labels, conf_fips=base.conf_fips, conf_zone=base.conf_zone)
    else:
        store = base
    weights = synthetic.make_weights(fips, P, seed=3)
    core = HeatAlertCore(store, weights, "cuda:0")
    n = 70
    idx = {k: v.astype(np.int32) for k, v in dict(loc=rng.integers(0, S, n), year=rng.integers(0, Y, n),
                                                    coef_col=rng.integers(0, S, n), coef_idx=rng.integers(0, P, n)).items()}
    state = core.alloc_state(n)
    obs = torch.zeros((n, 29), device="cuda")
    core.reset(state, obs, *(_dev(idx[k]) for k in ("loc", "year", "coef_col", "coef_idx")))
    st, o_ref = orc.batch_reset(store.feat, store.sig_code, store.n_days, idx["loc"], idx["year"], idx["coef_col"], idx["coef_idx"])
    assert_obs_equal(obs.cpu().numpy(), o_ref)
    reward = torch.zeros(n, device="cuda")
    term = torch.zeros(n, dtype=torch.uint8, device="cuda")
    for i in range(min(T + 3, 40) if T < 200 else T + 2):
        act = (rng.random(n) < 0.5).astype(np.uint8)
        core.step(state, _dev(act), obs, reward, term, variant=(1, 2, 3)[i % 3])
        o_ref, r_ref, d_ref, _, _ = orc.batch_step(store.feat, store.sig_code, weights.coef, st, act)
        assert_obs_equal(obs.cpu().numpy(), o_ref, f"step {i}")
        np.testing.assert_allclose(reward.cpu().numpy(), r_ref, rtol=REWARD_RTOL, atol=REWARD_ATOL)
        np.testing.assert_array_equal(term.cpu().numpy().astype(bool), d_ref)


@pytest.mark.parametrize("n_years,n_samples", [(18, 5), (9, 8), (25, 21)])
def test_table_variant_shapes_off_the_beaten_path(n_years, n_samples):
    """The TABLE variant's table build outside the real shapes: more than 16 years (its generic year-tile loop),
    fewer / a non-multiple of 8 posterior samples, coefficient columns whose county is not in the feature store,
    and the lock-step fast path itself (all envs canonical and on the same day) -- against the oracle."""
    from oracle import heat_alert_oracle as orc
    from weather2alert_b200 import synthetic
    from weather2alert_b200.core import HeatAlertCore

    fips, zones = synthetic.synthetic_counties(12, seed=n_years)
    years = list(range(2000, 2000 + n_years))
    store = synthetic.make_store(fips, years, seed=3, zones=zones)
    extra = [f for f in ("01001", "56045", "99999") if f not in fips]
    fips_w = sorted(fips + extra)  # columns of counties the store does not hold: skipped by the table build
    weights = synthetic.make_weights(fips_w, n_samples, seed=4)
    core = HeatAlertCore(store, weights, "cuda:0")
    assert (core.col_to_loc.cpu().numpy() < 0).sum() == len(extra)
    n = 1500
    rng = np.random.default_rng(11)
    loc = rng.integers(0, len(fips), n).astype(np.int32)
    col = np.array([fips_w.index(fips[i]) for i in loc], dtype=np.int32)  # canonical: the county's own column
    idx = dict(loc=loc, year=rng.integers(0, n_years, n).astype(np.int32), coef_col=col,
               coef_idx=rng.integers(0, n_samples, n).astype(np.int32), budget=rng.integers(0, 9, n).astype(np.int32))
    state = core.alloc_state(n)
    obs = torch.zeros((n, 29), device="cuda")
    core.reset(state, obs, *(_dev(idx[k]) for k in ("loc", "year", "coef_col", "coef_idx")), budget=_dev(idx["budget"]))
    assert bool((state.view(torch.int32).view(2, n, 2)[1, :, 1] < 0).all())  # every env carries the canonical bit
    st, o_ref = orc.batch_reset(store.feat, store.sig_code, store.n_days, idx["loc"], idx["year"], idx["coef_col"],
                                idx["coef_idx"], idx["budget"])
    reward = torch.zeros(n, device="cuda")
    term = torch.zeros(n, dtype=torch.uint8, device="cuda")
    for i in range(40):
        act = (rng.random(n) < 0.4).astype(np.int64)
        core.step(state, _dev(act), obs, reward, term, variant=VARIANTS["table"])
        o_ref, r_ref, d_ref, _, _ = orc.batch_step(store.feat, store.sig_code, weights.coef, st, act)
        assert_obs_equal(obs.cpu().numpy(), o_ref, f"step {i}")
        np.testing.assert_allclose(reward.cpu().numpy(), r_ref, rtol=REWARD_RTOL, atol=REWARD_ATOL)
        np.testing.assert_array_equal(term.cpu().numpy().astype(bool), d_ref)


# ---- the benchmarked shape itself (746 counties x 11 years x 153 days x 100 samples) ------------------------
@pytest.fixture(scope="module")
def real_shape():
    """The tables bench.py runs on, for both checkpoint layouts, and one core each (3.1 GB of tables per core)."""
    import bench
    from weather2alert_b200.core import HeatAlertCore

    out = {}
    for layout in ("linear", "nn"):
        store, weights = bench.make_tables(layout=layout)
        out[layout] = (store, weights, HeatAlertCore(store, weights, "cuda:0"))
    yield out
    out.clear()
    torch.cuda.empty_cache()


def _against_c_oracle(store, weights, core, n, steps, variants, seed, ragged=False):
    """n envs, env i on column i mod S_w of the weights and that county's weather (canonical), stepped by every
    variant and by oracle/c/heat_oracle.c from the same streams.  ragged: a third of the envs are reset mid-way,
    so the envs are on different days of the summer (the table variant looks every env's own day up)."""
    from oracle import c_oracle

    rng = np.random.default_rng(seed)
    Sw = len(weights.fips_list)
    col = (np.arange(n) % Sw).astype(np.int32)
    loc = np.array([store.fips_index[f] for f in weights.fips_list], np.int32)[col]
    year = rng.integers(0, store.shape[1], n).astype(np.int32)
    p = rng.integers(0, weights.n_samples, n).astype(np.int32)
    acts = [(rng.random(n) < (0.1, 0.5, 0.9)[i % 3]).astype(np.uint8) for i in range(steps)]
    co = c_oracle.COracle(store.feat, store.sig_code, store.n_days, weights.coef, n)
    ref = [(co.reset(loc, year, col, p).astype(np.float32),)]
    mask = (rng.random(n) < 0.33)
    year2, p2 = rng.integers(0, store.shape[1], n).astype(np.int32), rng.integers(0, weights.n_samples, n).astype(np.int32)
    for i, a in enumerate(acts):
        if ragged and i == steps // 2:  # the oracle has no masked reset: rebuild the masked envs' state by hand
            keep = {k: v.copy() for k, v in co.arr.items()}
            co.reset(np.where(mask, loc, loc), np.where(mask, year2, year), col, np.where(mask, p2, p))
            for k, v in co.arr.items():
                sel = mask if v.ndim == 1 else mask[:, None]
                v[...] = np.where(sel, v, keep[k])
        o, r, d, rem, atb = co.step(a)
        ref.append((o.astype(np.float32), r.copy(), d.copy(), rem.copy(), atb.copy()))
    results = {}
    for vname in variants:
        state = core.alloc_state(n)
        obs = torch.zeros((n, 29), device="cuda")
        reward, term = torch.zeros(n, device="cuda"), torch.zeros(n, dtype=torch.uint8, device="cuda")
        rem, atb = torch.zeros(n, dtype=torch.int32, device="cuda"), torch.zeros(n, dtype=torch.uint8, device="cuda")
        core.reset(state, obs, _dev(loc), _dev(year), _dev(col), _dev(p), remaining=rem)
        assert_obs_equal(obs.cpu().numpy(), ref[0][0], f"{vname} reset")
        worst = 0.0
        for i, a in enumerate(acts):
            if ragged and i == steps // 2:
                core.reset(state, obs, _dev(loc), _dev(np.where(mask, year2, year)), _dev(col), _dev(np.where(mask, p2, p)),
                           mask=_dev(mask.astype(np.uint8)), remaining=rem)
            core.step(state, _dev(a), obs, reward, term, rem, atb, variant=VARIANTS[vname])
            o, r, d, rm, ab = ref[i + 1]
            assert_obs_equal(obs.cpu().numpy(), o, f"{vname} step {i}")
            got = reward.cpu().numpy()
            np.testing.assert_allclose(got, r, rtol=REWARD_RTOL, atol=REWARD_ATOL, err_msg=f"{vname} step {i}")
            worst = max(worst, float((np.abs(got - r) / np.maximum(np.abs(r), 1e-5)).max()))
            np.testing.assert_array_equal(term.cpu().numpy(), d)
            np.testing.assert_array_equal(rem.cpu().numpy(), rm)
            np.testing.assert_array_equal(atb.cpu().numpy(), ab)
        results[vname] = worst
    return results


def test_real_shape_million_envs_vs_c_oracle(real_shape):
    """BASELINE configs[2] size on the bench's own tables: 2**20 envs x 20 steps, all three variants against the C
    oracle (VERDICT r1 'weak' item 1: the benchmarked shape compared with the oracle, not with itself)."""
    store, weights, core = real_shape["linear"]
    assert core.has_tables
    worst = _against_c_oracle(store, weights, core, 1 << 20, 20, ("table", "coop", "thread"), seed=7)
    assert max(worst.values()) < 1e-5


@pytest.mark.parametrize("n", [4096, 75_776, 75_809])
def test_real_shape_small_sets_both_table_instances(real_shape, n):
    """BASELINE configs[1] (4 096 envs) and the two sides of the switch between the latency-tuned instance of the TABLE
    step (one warp per CTA, every gather of a group in flight; up to 16 groups per SM = 75 776 envs on 148 SMs) and the
    bandwidth instance, on the real-shape tables against the C oracle."""
    store, weights, core = real_shape["linear"]
    worst = _against_c_oracle(store, weights, core, n, 10, ("table", "coop"), seed=n)
    assert max(worst.values()) < 1e-5


def test_table_instances_and_the_locality_hint_never_change_a_result(real_shape):
    """`W2A_STEP_TABLE` has three instances (one / two / eight row rounds in flight, picked by env-set size and by the
    `W2A_STEP_LOCAL` hint): same bytes out of each, on interleaved and on sorted slots."""
    import bench
    from weather2alert_b200 import _lib

    store, weights, core = real_shape["linear"]
    n = 200_000  # above the small-set switch, below the one-round limit: the hint switches one <-> two rounds
    for locality in ("interleaved", "sorted"):
        s = bench.index_streams(store, weights, n, 0, False, locality)
        idx = {k: torch.from_numpy(v).cuda() for k, v in s.items()}
        out = []
        for variant in (_lib.STEP_TABLE, _lib.STEP_TABLE | _lib.STEP_LOCAL):
            st, ob = core.alloc_state(n), torch.zeros((n, 29), device="cuda")
            core.reset(st, ob, idx["loc"], idx["year"], idx["coef_col"], idx["coef_idx"])
            rew, term = torch.zeros(n, device="cuda"), torch.zeros(n, dtype=torch.uint8, device="cuda")
            rem, atb = torch.zeros(n, dtype=torch.int32, device="cuda"), torch.zeros(n, dtype=torch.uint8, device="cuda")
            g = torch.Generator(device="cuda").manual_seed(5)
            for _ in range(6):
                a = (torch.rand(n, generator=g, device="cuda") < 0.4).to(torch.uint8)
                core.step(st, a, ob, rew, term, rem, atb, variant=variant)
            out.append((st.clone(), ob.clone(), rew.clone(), term.clone(), rem.clone(), atb.clone()))
        for x, y in zip(*out):
            assert torch.equal(x.view(torch.uint8), y.view(torch.uint8)), locality


def test_real_shape_ragged_days_on_the_table_path(real_shape):
    """Envs on different days of the summer (masked reset mid-way) keep using the tabulated logits: every env looks
    its own day up in the whole-summer tables; checked against the C oracle on the real shape."""
    store, weights, core = real_shape["linear"]
    _against_c_oracle(store, weights, core, 40_000, 30, ("table",), seed=9, ragged=True)


def test_nn_checkpoint_layout_720_counties(real_shape):
    """BASELINE configs[2] literally: the 720-county `fips_list` of weights/nn_full_medicare_all/config.yaml
    (tools/make_nn_fips.py) against a 746-county store -- coefficient column != store county, 26 counties of the
    store without a column; all variants against the C oracle, and augmentation on that list."""
    store, weights, core = real_shape["nn"]
    assert len(weights.fips_list) == 720 and store.shape[0] == 746
    c2l = core.col_to_loc.cpu().numpy()
    assert (c2l >= 0).all() and len(set(c2l.tolist())) == 720 and (c2l != np.arange(720)).any()
    _against_c_oracle(store, weights, core, 100_000, 12, ("table", "coop", "thread"), seed=11)
    # similar-county augmentation (Q6): column = index inside the similar list -> not canonical, direct path
    from oracle import heat_alert_oracle as orc
    from weather2alert_b200.sampling import EpisodeSampler

    n = 5000
    rng = np.random.default_rng(3)
    sp = EpisodeSampler(store, weights)
    s = sp.draw_bulk(rng, n, rng.integers(0, 720, n), True)
    state, obs = core.alloc_state(n), torch.zeros((n, 29), device="cuda")
    reward, term = torch.zeros(n, device="cuda"), torch.zeros(n, dtype=torch.uint8, device="cuda")
    core.reset(state, obs, *(_dev(s[k]) for k in ("loc", "year", "coef_col", "coef_idx")))
    st, o_ref = orc.batch_reset(store.feat, store.sig_code, store.n_days, s["loc"], s["year"], s["coef_col"], s["coef_idx"])
    assert_obs_equal(obs.cpu().numpy(), o_ref)
    for i in range(6):
        act = (rng.random(n) < 0.5).astype(np.uint8)
        core.step(state, _dev(act), obs, reward, term, variant=(3, 2, 1)[i % 3])
        o_ref, r_ref, d_ref, _, _ = orc.batch_step(store.feat, store.sig_code, weights.coef, st, act)
        assert_obs_equal(obs.cpu().numpy(), o_ref, f"step {i}")
        np.testing.assert_allclose(reward.cpu().numpy(), r_ref, rtol=REWARD_RTOL, atol=REWARD_ATOL)


def test_table_variant_needs_the_precomputed_tables(golden):
    """W2A_STEP_TABLE without w2a_precompute_logits is refused loudly; with it, a later w2a_set_* drops the tables."""
    import ctypes as C

    from weather2alert_b200 import _lib
    from weather2alert_b200.core import HeatAlertCore

    core = HeatAlertCore(golden.store(), golden.weights("std"), "cuda:0", precompute=False)
    assert not core.has_tables and core.table_bytes > 0
    n = 64
    state, obs = core.alloc_state(n), torch.zeros((n, 29), device="cuda")
    z = torch.zeros(n, dtype=torch.int32, device="cuda")
    core.reset(state, obs, z, z, z, z)
    reward, term = torch.zeros(n, device="cuda"), torch.zeros(n, dtype=torch.uint8, device="cuda")
    with pytest.raises(_lib.W2AError, match="w2a_precompute_logits"):
        core.step(state, z.to(torch.uint8), obs, reward, term, variant=_lib.STEP_TABLE)
    assert core.precompute() and core.has_tables
    core.step(state, z.to(torch.uint8), obs, reward, term, variant=_lib.STEP_TABLE)
    torch.cuda.synchronize()
    lib = _lib.lib()  # re-installing the coefficients invalidates the tables built from them
    _lib.check(lib.w2a_set_coefs(core.ctx, C.c_void_p(core.w.data_ptr()), core.P, core.Sw), core.ctx)
    with pytest.raises(_lib.W2AError):
        core.step(state, z.to(torch.uint8), obs, reward, term, variant=_lib.STEP_TABLE)


def test_table_variant_float32_logits_fall_back_to_float64_where_it_matters():
    """The step kernel gathers (z_b, z_e) as float32 from the one-sector table; where that rounding could reach the
    reward -- a huge tabulated logit that the agent terms pull back into the sigmoid's sensitive range -- it must
    take the float64 pair instead.  Coefficients built for exactly that: bias +-1000 against remaining_budget x -+20
    with budgets around 50, plus moderate and saturated cases; TABLE against the oracle and against THREAD."""
    from oracle import heat_alert_oracle as orc
    from weather2alert_b200 import schema, synthetic
    from weather2alert_b200.core import HeatAlertCore
    from weather2alert_b200.store import RewardWeights

    S, Y, P, n = 9, 3, 12, 6000
    rng = np.random.default_rng(77)
    fips, zones = synthetic.synthetic_counties(S, seed=5)
    store = synthetic.make_store(fips, [2006, 2007, 2008], seed=6, zones=zones)
    coef = synthetic.make_weights(fips, P, seed=7).coef.copy()  # [P, S, 2, 28]
    rb = schema.SLOT_REMAINING_BUDGET
    for p in range(P):  # sample p: bias = sign * 20 * B, remaining-budget coefficient = -sign * 20  ->  z crosses 0 near rem = B
        sign, B = (1.0 if p % 2 else -1.0), (48 + p)
        for h in (0, 1):
            coef[p, :, h, 27] += sign * 20.0 * B
            coef[p, :, h, rb] = -sign * 20.0
    coef[0] *= 0.01  # one moderate sample (|z| < 32 everywhere)
    weights = RewardWeights(fips_list=list(fips), coef=coef.astype(np.float32))
    core = HeatAlertCore(store, weights, "cuda:0")
    loc = rng.integers(0, S, n).astype(np.int32)
    idx = dict(loc=loc, year=rng.integers(0, Y, n).astype(np.int32), coef_col=loc.copy(),
               coef_idx=rng.integers(0, P, n).astype(np.int32), budget=rng.integers(45, 64, n).astype(np.int32))
    outs = {}
    for vname in ("table", "thread"):
        state, obs = core.alloc_state(n), torch.zeros((n, 29), device="cuda")
        core.reset(state, obs, *(_dev(idx[k]) for k in ("loc", "year", "coef_col", "coef_idx")), budget=_dev(idx["budget"]))
        st, _ = orc.batch_reset(store.feat, store.sig_code, store.n_days, idx["loc"], idx["year"], idx["coef_col"],
                                idx["coef_idx"], idx["budget"])
        reward, term = torch.zeros(n, device="cuda"), torch.zeros(n, dtype=torch.uint8, device="cuda")
        r2 = np.random.default_rng(1)
        rews, sensitive = [], 0
        for i in range(30):
            act = (r2.random(n) < 0.7).astype(np.uint8)
            core.step(state, _dev(act), obs, reward, term, variant=VARIANTS[vname])
            o_ref, r_ref, d_ref, _, _ = orc.batch_step(store.feat, store.sig_code, weights.coef, st, act)
            assert_obs_equal(obs.cpu().numpy(), o_ref, f"{vname} step {i}")
            got = reward.cpu().numpy()
            np.testing.assert_allclose(got, r_ref, rtol=REWARD_RTOL, atol=REWARD_ATOL, err_msg=f"{vname} step {i}")
            sensitive += int(((np.abs(r_ref) > 1e-3) & (np.abs(r_ref) < 6.0) & (idx["coef_idx"] != 0)).sum())
            rews.append(got)
        outs[vname] = np.stack(rews)
        assert sensitive > 1000  # rewards of the +-1000-bias samples away from both saturations: their logits did cross
    np.testing.assert_allclose(outs["table"], outs["thread"], rtol=3e-6, atol=REWARD_ATOL)
