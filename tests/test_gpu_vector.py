"""Vector env behaviour and size-independent properties at BASELINE.json sizes."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def big_core():
    """746 counties x 11 years, P=100: the shape of BASELINE.json configs[1..3]."""
    from weather2alert_b200 import synthetic
    from weather2alert_b200.core import HeatAlertCore

    fips, zones = synthetic.synthetic_counties(746, seed=0)
    store = synthetic.make_store(fips, list(range(2006, 2017)), seed=1, zones=zones)
    weights = synthetic.make_weights(fips, 100, seed=2)
    return HeatAlertCore(store, weights, "cuda:0")


def test_vector_env_autoreset_and_oracle(golden):
    from oracle import heat_alert_oracle as orc
    from weather2alert_b200 import HeatAlertVectorEnv

    store, weights = golden.store(), golden.weights("std")
    n = 300
    env = HeatAlertVectorEnv(n, weights=weights, store=store, seed=11, similar_climate_counties=True)
    obs, info = env.reset()
    assert obs.shape == (n, 29) and obs.is_cuda and obs.dtype == torch.float32
    idx = {k: v.cpu().numpy() for k, v in env._idx.items()}
    st, o_ref = orc.batch_reset(store.feat, store.sig_code, store.n_days, idx["loc"], idx["year"], idx["coef_col"],
                                idx["coef_idx"], idx.get("budget"))
    np.testing.assert_array_equal(obs.cpu().numpy(), o_ref.astype(np.float32))
    rng = np.random.default_rng(1)
    n_resets = 0
    for i in range(320):
        act = torch.from_numpy((rng.random(n) < 0.4)).cuda()
        obs, rew, term, trunc, info = env.step(act)
        assert term.dtype == torch.bool and not trunc.any()
        o_ref, r_ref, d_ref, rem_ref, atb_ref = orc.batch_step(store.feat, store.sig_code, weights.coef, st, act.cpu().numpy())
        np.testing.assert_array_equal(term.cpu().numpy(), d_ref)
        np.testing.assert_allclose(rew.cpu().numpy(), r_ref, rtol=1e-5, atol=1e-10)
        done = d_ref
        if done.any():  # SAME_STEP autoreset: those envs now show the first obs of a new episode
            n_resets += int(done.sum())
            idx = {k: v.cpu().numpy() for k, v in env._idx.items()}
            st2, o2 = orc.batch_reset(store.feat, store.sig_code, store.n_days, idx["loc"], idx["year"],
                                      idx["coef_col"], idx["coef_idx"], idx.get("budget"))
            for k in ("loc", "year", "col", "p", "budget", "t", "streak", "n_days", "obs"):
                st[k] = np.where(done.reshape((-1,) + (1,) * (st[k].ndim - 1)), st2[k], st[k])
            st["hist"][done] = 0
            o_ref = st["obs"]
        np.testing.assert_array_equal(obs.cpu().numpy()[~done], o_ref[~done].astype(np.float32))
        np.testing.assert_array_equal(obs.cpu().numpy()[done], st["obs"][done].astype(np.float32))
    assert n_resets >= 2 * n  # every env finished at least two summers (one of them is 150 days long)


def test_vector_env_state_dict_replay(golden):
    from weather2alert_b200 import HeatAlertVectorEnv

    env = HeatAlertVectorEnv(64, weights=golden.weights("std"), store=golden.store(), seed=5)
    env.reset()
    acts = [torch.from_numpy(np.random.default_rng(i).random(64) < 0.5).cuda() for i in range(30)]
    for a in acts[:10]:
        env.step(a)
    snap = env.state_dict()
    first = [tuple(x.clone() for x in env.step(a)[:3]) for a in acts[10:]]
    env.load_state_dict(snap)
    again = [tuple(x.clone() for x in env.step(a)[:3]) for a in acts[10:]]
    for f, g in zip(first, again):
        for x, y in zip(f, g):
            assert torch.equal(x.view(torch.uint8), y.view(torch.uint8))


def test_state_dict_resume_crosses_an_autoreset_with_reset_options(golden):
    """a FRESH env that loads the checkpoint must autoreset with the options of the original
    `reset` (similar-county augmentation, budget sampling) and show the same infos and statistics."""
    from weather2alert_b200 import HeatAlertVectorEnv

    n = 96
    kw = dict(weights=golden.weights("std"), store=golden.store(), track_stats=True)
    opts = dict(similar_climate_counties=True, budget=7, sample_budget=True, sample_budget_type="centered")
    env = HeatAlertVectorEnv(n, seed=5, **kw)
    env.reset(options=opts)
    acts = [torch.from_numpy(np.random.default_rng(i).random(n) < 0.5).cuda() for i in range(200)]
    for a in acts[:100]:
        env.step(a)
    snap = env.state_dict()

    def run(e):
        out = []
        for a in acts[100:]:  # crosses the end of the 150-day and of the 153-day summers
            o, r, t, _, info = e.step(a)
            out.append((o.clone(), r.clone(), t.clone(), info["remaining_budget"].clone(), info["at_budget"].clone(),
                        info["location_index"].clone()))
        return out, e.stats.totals.clone()

    first, tot1 = run(env)
    fresh = HeatAlertVectorEnv(n, seed=999, **kw)  # different seed, default options: everything comes from the checkpoint
    fresh.load_state_dict(snap)
    assert fresh._variant() == env._variant()
    i0 = fresh._infos()
    assert torch.equal(i0["remaining_budget"], snap["rem"]) and torch.equal(i0["at_budget"], snap["atb"])
    again, tot2 = run(fresh)
    for f, g in zip(first, again):
        for x, y in zip(f, g):
            assert torch.equal(x.view(torch.uint8), y.view(torch.uint8))
    assert torch.equal(tot1, tot2)


def test_autoreset_infos_and_masked_draws_stay_unsorted(golden):
    """SAME_STEP autoreset keeps the terminal step's info in `final_info`; a masked autoreset (ragged summers) does
    not sort its draws (sorting a subset would tie slots to county ranges), a full reset does."""
    from weather2alert_b200 import HeatAlertVectorEnv

    store, weights = golden.store(), golden.weights("std")
    n = 400
    env = HeatAlertVectorEnv(n, weights=weights, store=store, seed=3, budget=3)
    env.reset()
    key = lambda d: ((d["loc"].long() * 1000 + d["coef_col"]) * 1000 + d["coef_idx"]) * 100 + d["year"]  # noqa: E731
    k0 = key(env._idx)
    assert bool((k0[1:] >= k0[:-1]).all())  # the full reset is sorted
    saw_partial = False
    rng = np.random.default_rng(0)
    for i in range(160):
        prev_rem = env._rem.clone()
        act = torch.from_numpy(rng.random(n) < 0.2).cuda()
        obs, rew, term, _, info = env.step(act)
        if bool(term.any()):
            fi = info["final_info"]
            assert torch.equal(fi["_mask"], term)
            # the terminal step's remaining budget (post-action) is what final_info holds; infos show the new episode
            assert bool((fi["remaining_budget"][term] <= prev_rem[term]).all())
            assert bool((info["remaining_budget"][term] == 3).all()) and not bool(info["at_budget"][term].any())
            if not bool(term.all()):
                saw_partial = True
                k = key(env._idx)
                assert not bool((k[1:] >= k[:-1]).all())  # raw i.i.d. draws
            ep = info["episode_index"].numpy()
            lo, yr = env._ep_loc.cpu().numpy(), env._ep_year.cpu().numpy()
            assert ep[0] == f"{store.fips[lo[0]]}_{store.years[yr[0]]}"
        else:
            assert "final_info" not in info
    assert saw_partial


def test_million_envs_properties(big_core):
    """BASELINE configs[2] size (2**20 envs): the two step variants agree, the launch is
    deterministic, halves equal the whole (sharding), and the budget invariants hold."""
    core = big_core
    n = 1 << 20
    rng = np.random.default_rng(42)
    S, Y, P = core.S, core.Y, core.P
    mk = lambda a: torch.from_numpy(a.astype(np.int32)).cuda()  # noqa: E731
    loc = mk(np.arange(n) % S)
    year, cidx = mk(rng.integers(0, Y, n)), mk(rng.integers(0, P, n))
    budget = mk(rng.integers(0, 10, n))

    def run(variant, lo, hi, steps=20):
        m = hi - lo
        state = core.alloc_state(m)
        obs = torch.zeros((m, 29), device="cuda")
        core.reset(state, obs, loc[lo:hi].contiguous(), year[lo:hi].contiguous(), loc[lo:hi].contiguous(),
                   cidx[lo:hi].contiguous(), budget=budget[lo:hi].contiguous())
        reward = torch.zeros(m, device="cuda")
        term = torch.zeros(m, dtype=torch.uint8, device="cuda")
        rem = torch.zeros(m, dtype=torch.int32, device="cuda")
        atb = torch.zeros(m, dtype=torch.uint8, device="cuda")
        issued = torch.zeros(m, dtype=torch.int32, device="cuda")
        ret = torch.zeros(m, dtype=torch.float64, device="cuda")
        g = torch.Generator(device="cuda").manual_seed(7)
        for _ in range(steps):
            act = (torch.rand(n, device="cuda", generator=g) < 0.5)[lo:hi].to(torch.uint8)
            prev = rem.clone() if _ else budget[lo:hi].clone()
            core.step(state, act, obs, reward, term, rem, atb, variant=variant)
            issued += prev - rem
            ret += reward.double()
            assert int(rem.min()) >= 0                               # never over budget
            assert bool(((prev - rem == 1) | (prev - rem == 0)).all())
            assert bool((atb.bool() == (prev == 0)).all())           # at_budget is the pre-action flag
            assert bool((obs[:, 25] == rem.float()).all()) and bool((obs[:, 28] <= 14).all())
            assert bool((reward <= 0).all()) and bool(torch.isfinite(reward).all()) and not term.any()
        assert bool((issued <= budget[lo:hi]).all())
        return obs, ret, rem, state

    o1, r1, m1, s1 = run(1, 0, n)
    o2, r2, m2, s2 = run(2, 0, n)
    ot, rt, mt, st_ = run(3, 0, n)  # per-step logit table variant (all envs canonical and in lock step)
    assert torch.equal(ot.view(torch.int32), o2.view(torch.int32)) and torch.equal(mt, m2) and torch.equal(st_, s2)
    # the table variant gathers (z_b, z_e) as float32 (each exact to 1e-6 below |z| = 32): sums of rewards within 3e-6
    assert torch.allclose(rt, r2, rtol=3e-6, atol=1e-9)
    assert torch.equal(o1.view(torch.int32), o2.view(torch.int32)) and torch.equal(m1, m2) and torch.equal(s1, s2)
    assert torch.allclose(r1, r2, rtol=1e-6, atol=1e-9)  # float64 sums of float32 rewards; only summation order differs
    o3, r3, m3, s3 = run(2, 0, n)
    assert torch.equal(r2, r3) and torch.equal(o2.view(torch.int32), o3.view(torch.int32))  # deterministic
    half = n // 2
    oa, ra, _, _ = run(2, 0, half)
    ob, rb, _, _ = run(2, half, n)
    assert torch.equal(torch.cat([oa, ob]).view(torch.int32), o2.view(torch.int32))  # shards == whole
    assert torch.equal(torch.cat([ra, rb]), r2)


def test_vector_env_rollout_stats(golden):
    """track_stats=True accumulates the callback statistics on the device (callbacks.py:63-74)."""
    from weather2alert_b200 import HeatAlertVectorEnv
    from weather2alert_b200.distributed import summarize

    n = 128
    env = HeatAlertVectorEnv(n, weights=golden.weights("std"), store=golden.store(), seed=3, years=[2006],
                             track_stats=True)
    env.reset()
    rng = np.random.default_rng(0)
    ret = np.zeros(n)
    finished, alerts, refused = [], 0, 0
    for _ in range(153 * 2):
        a = rng.random(n) < 0.3
        atb_before = env._rem.cpu().numpy() == 0
        obs, rew, term, _, info = env.step(torch.from_numpy(a).cuda())
        ret += rew.cpu().numpy().astype(np.float64)
        alerts += int((a & ~atb_before).sum())
        refused += int((a & atb_before).sum())
        d = term.cpu().numpy()
        finished += list(ret[d])
        ret[d] = 0
    s = summarize(env.stats.reduce())
    assert s["episodes"] == len(finished) == 2 * n and s["steps"] == 153 * 2 * n
    assert s["alerts"] == alerts and s["refused_at_budget"] == refused
    assert s["mean_return"] == pytest.approx(np.mean(finished), rel=1e-9)
    assert s["std_return"] == pytest.approx(np.std(finished), rel=1e-6)


def test_sort_episodes_is_a_permutation_of_the_same_draws(golden):
    from weather2alert_b200 import HeatAlertVectorEnv

    kw = dict(weights=golden.weights("std"), store=golden.store(), seed=9, similar_climate_counties=True)
    a = HeatAlertVectorEnv(500, sort_episodes=False, **kw)
    b = HeatAlertVectorEnv(500, **kw)  # the default for device-sampled resets of the whole batch
    assert b.sort_episodes and not a.sort_episodes
    a.reset()
    b.reset()
    ka = torch.stack([a._idx[k] for k in ("loc", "coef_col", "coef_idx", "year")], 1).cpu().numpy()
    kb = torch.stack([b._idx[k] for k in ("loc", "coef_col", "coef_idx", "year")], 1).cpu().numpy()
    assert sorted(map(tuple, ka)) == list(map(tuple, kb))  # same multiset, lexicographic order
    assert torch.equal(b._list_idx, b._idx["list_idx"])


def test_graphed_policy_rollout_equals_eager_loop():
    """policy(obs) -> step, 153 times, captured in one CUDA graph == the same loop launched eagerly."""
    from weather2alert_b200 import HeatAlertVectorEnv, synthetic

    fips, zones = synthetic.synthetic_counties(20, seed=2)
    store = synthetic.make_store(fips, [2006, 2007], seed=3, zones=zones)
    weights = synthetic.make_weights(fips, 8, seed=4)
    n = 3000
    w_pol = torch.randn(29, device="cuda") * 0.5

    def policy(obs):  # deterministic logistic policy on the observation (NaN significance -> 0)
        return (torch.nan_to_num(obs) @ w_pol) > 0.3

    kw = dict(weights=weights, store=store, seed=12, budget=6)
    eager = HeatAlertVectorEnv(n, autoreset=False, **kw)
    obs, _ = eager.reset()
    ret = torch.zeros(n, dtype=torch.float64, device="cuda")
    for _ in range(153):
        obs, r, term, _, _ = eager.step(policy(obs))
        ret += r
    assert bool(term.all())

    graphed = HeatAlertVectorEnv(n, **kw)
    graphed.reset()
    roll = graphed.graphed_rollout(policy, reset_at_end=True)
    got = roll.replay().clone()
    assert torch.equal(got, ret)
    # the captured reset started a new summer: the first observation is a fresh day-0 row
    assert bool((graphed._obs[:, 17] == 0).all()) and graphed._step_no == 153
    second = roll.replay().clone()
    assert not torch.equal(second, got) and bool(torch.isfinite(second).all())
    with pytest.raises(ValueError):
        graphed.graphed_rollout(policy, steps=154)


def test_reference_rng_reset_equals_independent_single_envs(golden):
    """reset(seed=s, options={'reference_rng': True}): env i == HeatAlertEnv().reset(seed=s+i) -- the
    seeding rule of gymnasium.vector.SyncVectorEnv -- including augmentation and budget sampling."""
    from weather2alert_b200 import HeatAlertEnv, HeatAlertVectorEnv

    store, weights = golden.store(), golden.weights("std")
    n, seed = 12, 4321
    opts = dict(reference_rng=True, similar_climate_counties=True, sample_budget=True, sample_budget_type="centered")
    venv = HeatAlertVectorEnv(n, weights=weights, store=store, autoreset=False, budget=6)
    obs, info = venv.reset(seed=seed, options=opts)
    singles = []
    for i in range(n):
        e = HeatAlertEnv(weights=weights, store=store, budget=6)
        o, inf = e.reset(seed=seed + i, similar_climate_counties=True, sample_budget=True, sample_budget_type="centered")
        np.testing.assert_array_equal(obs[i].cpu().numpy().view(np.int32), o.view(np.int32))
        assert inf["remaining_budget"] == int(info["remaining_budget"][i])
        # the reference's info (env.py:228-236): location_index is the coefficient column (Q6: the index inside
        # the similar-county list), location the similar county, episode_index the ORIGINAL county's episode
        assert inf["location_index"] == int(info["location_index"][i])
        assert weights.fips_list.index(str(inf["location"])) == int(info["location_list_index"][i])
        assert str(info["location"][i]) == str(inf["location"]) and info["episode_index"][i] == inf["episode_index"]
        singles.append(e)
    assert len(info["location"]) == n and info["feature_names"] == singles[0].feat_names
    rng = np.random.default_rng(0)
    for _ in range(25):
        a = rng.integers(0, 2, n)
        obs, rew, term, _, info = venv.step(torch.from_numpy(a).cuda())
        for i, e in enumerate(singles):
            o, r, d, _, inf = e.step(int(a[i]))
            np.testing.assert_array_equal(obs[i].cpu().numpy().view(np.int32), o.view(np.int32))
            assert r == float(rew[i]) and d == bool(term[i]) and inf["at_budget"] == bool(info["at_budget"][i])


def test_vector_env_auto_variant_picks_table(golden, monkeypatch):
    """AUTO above the C ABI: the per-step logit table for large batches of envs on their own county's
    coefficients, COOP/THREAD otherwise -- and the choice never changes a result (ragged summers make
    the envs fall out of lock step after the first autoreset: those take the table kernel's direct path)."""
    from weather2alert_b200 import HeatAlertVectorEnv, _lib
    from weather2alert_b200.core import HeatAlertCore

    store, weights = golden.store(), golden.weights("std")
    n = 257
    small = HeatAlertVectorEnv(n, weights=weights, store=store, seed=5)
    small.reset()
    monkeypatch.setattr(HeatAlertCore, "TABLE_MIN_ENVS", 1000)
    assert small._variant() == _lib.STEP_AUTO  # below TABLE_MIN_ENVS
    monkeypatch.setattr(HeatAlertCore, "TABLE_MIN_ENVS", 1)
    assert small._variant() & 0xFF == _lib.STEP_TABLE
    aug = HeatAlertVectorEnv(n, weights=weights, store=store, seed=5, similar_climate_counties=True)
    aug.reset()
    assert aug._variant() == _lib.STEP_AUTO  # similar-county substitution: not the county's own column
    forced = HeatAlertVectorEnv(n, weights=weights, store=store, seed=5, step_variant=_lib.STEP_COOP)
    forced.reset()
    rng = np.random.default_rng(3)
    for i in range(200):
        act = torch.from_numpy(rng.random(n) < 0.3).cuda()
        o1, r1, t1, _, i1 = small.step(act)
        o2, r2, t2, _, i2 = forced.step(act)
        assert torch.equal(o1.view(torch.int32), o2.view(torch.int32)) and torch.equal(t1, t2), f"step {i}"
        assert torch.allclose(r1, r2, rtol=3e-6, atol=1e-10)  # table: float32 tabulated logits, see DESIGN.md
    assert (_lib.STEP_TABLE in {k[1] for k in small._bound}) and (_lib.STEP_COOP in {k[1] for k in forced._bound})


def test_vector_env_intended_augmentation(golden, monkeypatch):
    """`similar_climate_counties="intended"`: episodes move to a similar county (weather and column alike), stay
    canonical (so the TABLE variant applies) and step like any other episode."""
    from oracle import heat_alert_oracle as orc
    from weather2alert_b200 import HeatAlertVectorEnv, _lib
    from weather2alert_b200.core import HeatAlertCore

    store, weights = golden.store(), golden.weights("std")
    monkeypatch.setattr(HeatAlertCore, "TABLE_MIN_ENVS", 1)
    n = 500
    env = HeatAlertVectorEnv(n, weights=weights, store=store, seed=21, autoreset=False)
    obs, _ = env.reset(options={"similar_climate_counties": "intended"})
    idx = {k: v.cpu().numpy() for k, v in env._idx.items()}
    np.testing.assert_array_equal(idx["coef_col"], idx["list_idx"])
    np.testing.assert_array_equal(idx["loc"], env.sampler.list_to_data[idx["list_idx"]])
    assert env._variant() & 0xFF == _lib.STEP_TABLE
    with pytest.raises(ValueError):
        env.reset(seed=1, options={"similar_climate_counties": "intended", "reference_rng": True})
    obs, _ = env.reset(options={"similar_climate_counties": "intended"})
    idx = {k: v.cpu().numpy() for k, v in env._idx.items()}
    st, o_ref = orc.batch_reset(store.feat, store.sig_code, store.n_days, idx["loc"], idx["year"], idx["coef_col"],
                                idx["coef_idx"], idx.get("budget"))
    np.testing.assert_array_equal(obs.cpu().numpy(), o_ref.astype(np.float32))
    rng = np.random.default_rng(2)
    for _ in range(12):
        act = rng.random(n) < 0.5
        obs, rew, term, _, _ = env.step(torch.from_numpy(act).cuda())
        o_ref, r_ref, d_ref, _, _ = orc.batch_step(store.feat, store.sig_code, weights.coef, st, act)
        np.testing.assert_array_equal(obs.cpu().numpy(), o_ref.astype(np.float32))
        np.testing.assert_allclose(rew.cpu().numpy(), r_ref, rtol=1e-5, atol=1e-10)


def test_index_streams_pick_the_table_variant_when_every_env_is_on_its_own_column(golden, monkeypatch):
    """Caller-supplied index streams: the vector env checks on the host whether every (county, column) pair is the
    county's own and only then steps on the tabulated logits; results equal the cooperative kernel's either way."""
    from weather2alert_b200 import HeatAlertVectorEnv, _lib
    from weather2alert_b200.core import HeatAlertCore

    monkeypatch.setattr(HeatAlertCore, "TABLE_MIN_ENVS", 1)
    store, weights = golden.store(), golden.weights("std")
    n = 300
    rng = np.random.default_rng(4)
    col = rng.integers(0, len(weights.fips_list), n)
    loc = np.array([store.fips_index[weights.fips_list[c]] for c in col])
    streams = dict(loc=loc, year=rng.integers(0, store.shape[1], n), coef_col=col, coef_idx=rng.integers(0, weights.n_samples, n))
    streams["year"] = np.where(store.n_days[loc, streams["year"]] == 0, 0, streams["year"])
    a = HeatAlertVectorEnv(n, weights=weights, store=store, autoreset=False)
    a.reset(options={"index_streams": streams})
    assert a._variant() & 0xFF == _lib.STEP_TABLE
    mixed = dict(streams, coef_col=np.roll(col, 1))  # somebody else's column: not tabulated
    b = HeatAlertVectorEnv(n, weights=weights, store=store, autoreset=False)
    b.reset(options={"index_streams": mixed})
    assert b._variant() == _lib.STEP_AUTO
    c = HeatAlertVectorEnv(n, weights=weights, store=store, autoreset=False, step_variant=_lib.STEP_COOP)
    c.reset(options={"index_streams": streams})
    for i in range(30):
        act = torch.from_numpy(rng.random(n) < 0.4).cuda()
        oa, ra, ta, _, _ = a.step(act)
        oc, rc, tc, _, _ = c.step(act)
        assert torch.equal(oa.view(torch.int32), oc.view(torch.int32)) and torch.equal(ta, tc)
        assert torch.allclose(ra, rc, rtol=3e-6, atol=1e-10)
