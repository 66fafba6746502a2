"""The episode-fused rollout kernel against step-by-step execution and against the oracle."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _case(seed, n, short=True):
    from weather2alert_b200 import synthetic

    rng = np.random.default_rng(seed)
    S, Y, P = 30, 3, 16
    fips, zones = synthetic.synthetic_counties(S, seed=seed)
    n_days = np.full((S, Y), 153, np.uint8)
    if short:
        n_days[rng.integers(0, S, 6), rng.integers(0, Y, 6)] = rng.integers(1, 152, 6)
    store = synthetic.make_store(fips, [2006, 2007, 2008], seed=seed + 1, zones=zones, n_days=n_days, p_alert=0.08)
    weights = synthetic.make_weights(fips, P, seed=seed + 2)
    idx = dict(loc=rng.integers(0, S, n), year=rng.integers(0, Y, n), coef_col=rng.integers(0, S, n),
               coef_idx=rng.integers(0, P, n), budget=rng.integers(-1, 12, n))
    return store, weights, {k: v.astype(np.int32) for k, v in idx.items()}


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("n", [1, 3, 130, 4099])
def test_rollout_equals_stepping_and_oracle(n):
    """Action bank: chunks of 61 + 61 + 61 steps (crossing chunk-of-8 boundaries and episode ends)
    must leave the same state as 183 `step` launches, and the same returns as the oracle."""
    from oracle import heat_alert_oracle as orc
    from weather2alert_b200.core import HeatAlertCore
    from weather2alert_b200.distributed import STAT_NAMES

    store, weights, idx = _case(100 + n, n)
    core = HeatAlertCore(store, weights, "cuda:0")
    d = {k: _dev(v) for k, v in idx.items()}
    rng = np.random.default_rng(9)
    total = 183
    acts = (rng.random((total, n)) < 0.35).astype(np.uint8)

    def fresh():
        st = core.alloc_state(n)
        obs = torch.zeros((n, 29), device="cuda")
        core.reset(st, obs, d["loc"], d["year"], d["coef_col"], d["coef_idx"], budget=d["budget"])
        return st, obs

    # (1) fused rollout in three calls
    st_r, _ = fresh()
    ret = torch.zeros(n, dtype=torch.float64, device="cuda")
    stats = torch.zeros(15, dtype=torch.float64, device="cuda")
    for c in range(3):
        core.rollout(st_r, 61, ret, actions=_dev(acts[61 * c:61 * (c + 1)]), stats=stats)
    # (2) oracle, stopping every env at its terminal step
    ost, _ = orc.batch_reset(store.feat, store.sig_code, store.n_days, idx["loc"], idx["year"], idx["coef_col"],
                             idx["coef_idx"], idx["budget"])
    ret_ref = np.zeros(n)
    alive = np.ones(n, bool)
    n_steps = alerts = refused = 0
    issued_days = [[] for _ in range(n)]   # day of summer of every issued alert, per env
    streak_lens = []                        # lengths of streaks of issued alerts that ended
    # (3) the step kernel, frozen per env at its terminal step
    st_s, obs = fresh()
    reward = torch.zeros(n, device="cuda")
    term = torch.zeros(n, dtype=torch.uint8, device="cuda")
    frozen = st_s.clone().view(2, n, 8)
    for k in range(total):
        before = ost["budget"] - ost["hist"].sum(axis=1)
        t_before, streak_before = ost["t"].copy(), ost["streak"].copy()
        _, r, dn, rem, atb = orc.batch_step(store.feat, store.sig_code, weights.coef, ost, acts[k])
        act_k = before - rem
        for i in np.flatnonzero(alive):
            if act_k[i]:
                issued_days[i].append(int(t_before[i]))
            elif streak_before[i] > 0:
                streak_lens.append(int(streak_before[i]))
        ret_ref += np.where(alive, r, 0.0)
        n_steps += int(alive.sum())
        alerts += int(((before - rem) * alive).sum())
        refused += int(((acts[k] != 0) & atb & alive).sum())
        core.step(st_s, _dev(acts[k]), obs, reward, term, variant=2)
        cur = st_s.view(2, n, 8)
        al = torch.from_numpy(alive).cuda()
        frozen[:, al] = cur[:, al]
        alive &= ~dn
    np.testing.assert_allclose(ret.cpu().numpy(), ret_ref, rtol=1e-9, atol=1e-12)
    assert torch.equal(st_r.view(2, n, 8), frozen)
    s = dict(zip(STAT_NAMES, stats.cpu().tolist()))
    finished = ~alive
    assert s["episodes"] == finished.sum() and s["steps"] == n_steps
    assert s["alerts"] == alerts and s["refused_at_budget"] == refused
    assert s["return_sum"] == pytest.approx(ret_ref[finished].sum(), rel=1e-9)
    assert s["return_sq_sum"] == pytest.approx((ret_ref[finished] ** 2).sum(), rel=1e-9)
    # timing statistics (callbacks.py:63-74 on issued alerts)
    all_days = np.array([d for days in issued_days for d in days], dtype=np.float64)
    assert s["alert_day_sum"] == all_days.sum() and s["alert_day_sq_sum"] == (all_days ** 2).sum()
    sl = np.array(streak_lens, dtype=np.float64)
    assert s["streaks"] == len(sl) and s["streak_sum"] == sl.sum() and s["streak_sq_sum"] == (sl ** 2).sum()
    # quantile days: only summers that ran from day 0 inside ONE call, i.e. those of at most 61 days here
    a50 = a80 = a100 = cnt = 0
    nd = store.n_days[idx["loc"], idx["year"]]
    for i in range(n):
        if finished[i] and nd[i] <= 61 and issued_days[i]:
            d = np.array(issued_days[i])
            tot = len(d)
            cum = np.arange(1, tot + 1)
            a50 += d[np.argmax(2 * cum >= tot)]
            a80 += d[np.argmax(5 * cum >= 4 * tot)]
            a100 += d[-1]
            cnt += 1
    assert (s["a50_sum"], s["a80_sum"], s["a100_sum"], s["a_count"]) == (a50, a80, a100, cnt)
    # a finished env is left alone by further rollouts
    snap = st_r.clone()
    core.rollout(st_r, 10, ret, actions=_dev(acts[:10]), stats=stats)
    assert torch.equal(st_r.view(2, n, 8)[:, torch.from_numpy(finished).cuda()],
                       snap.view(2, n, 8)[:, torch.from_numpy(finished).cuda()])


def test_rollout_hash_actions_match_host_stream(golden):
    """actions=None: the in-kernel Bernoulli stream equals `rollout_action_bits` on the host."""
    from weather2alert_b200 import HeatAlertVectorEnv
    from weather2alert_b200.distributed import rollout_action_bits

    n = 257
    env_a = HeatAlertVectorEnv(n, weights=golden.weights("std"), store=golden.store(), seed=5, years=[2006], autoreset=False)
    env_b = HeatAlertVectorEnv(n, weights=golden.weights("std"), store=golden.store(), seed=5, years=[2006], autoreset=False)
    env_a.reset()
    env_b.reset()
    ret_a, stats_a = env_a.rollout(p_alert=0.3, seed=77)
    bits = rollout_action_bits(77, np.arange(n), np.arange(153), 0.3)
    assert 0.25 < bits.mean() < 0.35
    ret_b, stats_b = env_b.rollout(actions=torch.from_numpy(bits))
    assert torch.equal(ret_a, ret_b)
    # counts are exact; the float sums go through atomics whose order is not fixed
    exact = [0, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14]
    assert torch.equal(stats_a[exact], stats_b[exact]) and float(stats_a[14]) > 0
    assert torch.allclose(stats_a, stats_b, rtol=1e-12, atol=0)
    assert torch.equal(env_a._state, env_b._state)
    assert float(stats_a[0]) == n and float(stats_a[3]) == 153 * n


def test_rollout_at_baseline_size_matches_stepping():
    """2**20 envs (BASELINE configs[2] size), a full summer: the fused rollout leaves the same state
    bytes as 153 step launches and its float64 returns equal the sum of the stepped float32 rewards."""
    from weather2alert_b200 import synthetic
    from weather2alert_b200.core import HeatAlertCore

    n, S, Y, P = 1 << 20, 64, 4, 32
    fips, zones = synthetic.synthetic_counties(S, seed=5)
    store = synthetic.make_store(fips, list(range(2006, 2006 + Y)), seed=6, zones=zones)
    core = HeatAlertCore(store, synthetic.make_weights(fips, P, seed=7), "cuda:0")
    g = torch.Generator(device="cuda").manual_seed(3)
    ri = lambda hi: torch.randint(0, hi, (n,), generator=g, device="cuda", dtype=torch.int32)  # noqa: E731
    loc, year, cidx, budget = ri(S), ri(Y), ri(P), ri(12)
    acts = (torch.rand((153, n), generator=g, device="cuda") < 0.3).to(torch.uint8)

    def fresh():
        st = core.alloc_state(n)
        obs = torch.zeros((n, 29), device="cuda")
        core.reset(st, obs, loc, year, loc, cidx, budget=budget)
        return st, obs

    st_r, _ = fresh()
    ret = torch.zeros(n, dtype=torch.float64, device="cuda")
    stats = torch.zeros(15, dtype=torch.float64, device="cuda")
    core.rollout(st_r, 153, ret, actions=acts, stats=stats)
    st_s, obs = fresh()
    reward = torch.zeros(n, device="cuda")
    term = torch.zeros(n, dtype=torch.uint8, device="cuda")
    total = torch.zeros(n, dtype=torch.float64, device="cuda")
    for k in range(153):
        core.step(st_s, acts[k], obs, reward, term)
        total += reward.double()
    assert bool(term.all()) and torch.equal(st_r, st_s)
    assert torch.allclose(ret, total, rtol=2e-6, atol=0)  # float32 rounding and the float32 sigmoid (<= 3e-7) of each stepped reward
    assert float(stats[0]) == n and float(stats[3]) == 153 * n
    assert float(stats[1]) == pytest.approx(float(ret.sum()), rel=1e-9)


@pytest.mark.parametrize("stochastic", [False, True])
@pytest.mark.parametrize("n", [1, 5, 1031])
def test_policy_rollout_matches_the_oracle_loop(n, stochastic):
    """`w2a_rollout_policy`: the agent inside the launch.  The oracle plays `a = policy(obs); obs, r = step(a)`
    (env.py:273-276 with a logistic policy on the float64 observation, significance slot skipped); returns,
    final state and the callback counts must agree."""
    from oracle import heat_alert_oracle as orc
    from weather2alert_b200.core import HeatAlertCore
    from weather2alert_b200.distributed import STAT_NAMES, rollout_hash, rollout_threshold

    store, weights, idx = _case(300 + n, n)
    core = HeatAlertCore(store, weights, "cuda:0")
    d = {k: _dev(v) for k, v in idx.items()}
    rng = np.random.default_rng(17 + n)
    theta = rng.normal(0.0, 0.4, 30)
    theta[6] = 0.0       # dos runs to 152: keep the logit in a range where both actions occur
    theta[25] = 0.15     # remaining budget
    theta[27] = 123.0    # the significance slot must be ignored
    theta[29] = -0.5
    use = np.array([s != 27 for s in range(29)])
    seed = 4242

    st = core.alloc_state(n)
    obs0 = torch.zeros((n, 29), device="cuda")
    core.reset(st, obs0, d["loc"], d["year"], d["coef_col"], d["coef_idx"], budget=d["budget"])
    ret = torch.zeros(n, dtype=torch.float64, device="cuda")
    stats = torch.zeros(15, dtype=torch.float64, device="cuda")
    core.rollout_policy(st, 153, ret, _dev(theta), stochastic=stochastic, seed=seed, stats=stats)

    ost, obs = orc.batch_reset(store.feat, store.sig_code, store.n_days, idx["loc"], idx["year"], idx["coef_col"],
                               idx["coef_idx"], idx["budget"])
    obs = obs.copy()
    ret_ref = np.zeros(n)
    alive = np.ones(n, bool)
    n_steps = alerts = refused = 0
    acts_played = np.zeros((153, n), np.uint8)
    for k in range(153):
        logit = obs[:, use] @ theta[:29][use] + theta[29]
        if stochastic:
            p = 1.0 / (1.0 + np.exp(-logit))
            a = rollout_hash(seed, np.arange(n), [k])[0] < rollout_threshold(p)
        else:
            a = logit > 0.0
        acts_played[k] = a
        before = ost["budget"] - ost["hist"].sum(axis=1)
        o, r, dn, rem, atb = orc.batch_step(store.feat, store.sig_code, weights.coef, ost, a.astype(np.uint8))
        ret_ref += np.where(alive, r, 0.0)
        n_steps += int(alive.sum())
        alerts += int(((before - rem) * alive).sum())
        refused += int((a & atb & alive).sum())
        obs = np.where((alive & ~dn)[:, None], o, obs)  # on and after the last day the old observation stands
        alive &= ~dn
    assert not alive.any()
    np.testing.assert_allclose(ret.cpu().numpy(), ret_ref, rtol=1e-9, atol=1e-12)
    s = dict(zip(STAT_NAMES, stats.cpu().tolist()))
    assert s["episodes"] == n and s["steps"] == n_steps
    assert s["alerts"] == alerts and s["refused_at_budget"] == refused
    if n > 100:
        assert 0 < alerts < n_steps and refused > 0  # the policy really decides both ways, and runs into budgets
    # the state the launch leaves = the state after stepping the same actions
    st2 = core.alloc_state(n)
    obs2 = torch.zeros((n, 29), device="cuda")
    core.reset(st2, obs2, d["loc"], d["year"], d["coef_col"], d["coef_idx"], budget=d["budget"])
    ret2 = torch.zeros(n, dtype=torch.float64, device="cuda")
    core.rollout(st2, 153, ret2, actions=_dev(acts_played))
    assert torch.equal(st, st2)
    # mid-summer envs are left alone
    st3 = core.alloc_state(n)
    core.reset(st3, obs2, d["loc"], d["year"], d["coef_col"], d["coef_idx"], budget=d["budget"])
    core.rollout(st3, 3, ret2, actions=_dev(acts_played[:3]))
    keep = st3.clone()
    r3 = torch.zeros(n, dtype=torch.float64, device="cuda")
    core.rollout_policy(st3, 153, r3, _dev(theta), stochastic=stochastic, seed=seed)
    mid = torch.from_numpy(store.n_days[idx["loc"], idx["year"]] > 3).cuda()
    assert torch.equal(st3.view(2, n, 8)[:, mid], keep.view(2, n, 8)[:, mid]) and not r3[mid].any()


@pytest.mark.parametrize("T", [200, 255])
def test_rollout_long_summers_quantile_days(T):
    """summers longer than 192 days (the C ABI allows T <= 255) used to overrun the
    rollout kernel's shared issued-alert bits.  A synthetic T-day store, one launch for the whole summer:
    returns against the oracle's loop and the a50/a80/a100 quantile days against a host recount."""
    from oracle import heat_alert_oracle as orc
    from weather2alert_b200 import schema, synthetic
    from weather2alert_b200.core import HeatAlertCore
    from weather2alert_b200.distributed import STAT_NAMES
    from weather2alert_b200.store import FeatureStore

    S, Y, P, n = 6, 2, 4, 333
    rng = np.random.default_rng(T)
    fips, zones = synthetic.synthetic_counties(S, seed=3)
    base = synthetic.make_store(fips, [2006, 2007], seed=4, zones=zones)
    reps = -(-T // schema.MAX_DAYS)
    feat = np.tile(base.feat, (1, 1, reps, 1))[:, :, :T].copy()
    sig = np.tile(base.sig_code, (1, 1, reps))[:, :, :T].copy()
    n_days = np.full((S, Y), T, np.uint8)
    n_days[0, 0] = 193  # just past the old 192-day limit
    store = FeatureStore(fips=base.fips, years=base.years, n_days=n_days, feat=feat, sig_code=sig,
                         sig_labels=base.sig_labels, conf_fips=base.conf_fips, conf_zone=base.conf_zone)
    weights = synthetic.make_weights(fips, P, seed=5)
    core = HeatAlertCore(store, weights, "cuda:0")
    idx = {k: v.astype(np.int32) for k, v in dict(loc=rng.integers(0, S, n), year=rng.integers(0, Y, n),
                                                    coef_col=rng.integers(0, S, n), coef_idx=rng.integers(0, P, n),
                                                    budget=rng.integers(20, 60, n)).items()}
    acts = (rng.random((T, n)) < 0.3).astype(np.uint8)
    acts[-30:] |= (rng.random((30, n)) < 0.5).astype(np.uint8)  # alerts late in the summer: days past 192 matter
    st = core.alloc_state(n)
    obs = torch.zeros((n, 29), device="cuda")
    core.reset(st, obs, *(_dev(idx[k]) for k in ("loc", "year", "coef_col", "coef_idx")), budget=_dev(idx["budget"]))
    ret = torch.zeros(n, dtype=torch.float64, device="cuda")
    stats = torch.zeros(15, dtype=torch.float64, device="cuda")
    core.rollout(st, T, ret, actions=_dev(acts), stats=stats)
    ost, _ = orc.batch_reset(store.feat, store.sig_code, store.n_days, idx["loc"], idx["year"], idx["coef_col"],
                             idx["coef_idx"], idx["budget"])
    ret_ref, alive = np.zeros(n), np.ones(n, bool)
    issued = [[] for _ in range(n)]
    for k in range(T):
        before = ost["budget"] - ost["hist"].sum(axis=1)
        t_before = ost["t"].copy()
        _, r, dn, rem, _ = orc.batch_step(store.feat, store.sig_code, weights.coef, ost, acts[k])
        for i in np.flatnonzero(alive & ((before - rem) > 0)):
            issued[i].append(int(t_before[i]))
        ret_ref += np.where(alive, r, 0.0)
        alive &= ~dn
    assert not alive.any()
    np.testing.assert_allclose(ret.cpu().numpy(), ret_ref, rtol=1e-9, atol=1e-12)
    s = dict(zip(STAT_NAMES, stats.cpu().tolist()))
    a50 = a80 = a100 = cnt = 0
    for d in issued:
        if d:
            d = np.array(d)
            cum = np.arange(1, len(d) + 1)
            a50 += d[np.argmax(2 * cum >= len(d))]
            a80 += d[np.argmax(5 * cum >= 4 * len(d))]
            a100 += d[-1]
            cnt += 1
    assert max(max(d) for d in issued if d) > 192
    assert (s["a50_sum"], s["a80_sum"], s["a100_sum"], s["a_count"]) == (a50, a80, a100, cnt)
    assert s["episodes"] == n
