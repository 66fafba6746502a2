"""The fused MLP policy kernel (`w2a_policy_mlp`, tensor cores, float16 inputs) against a plain PyTorch float32 reference of the
same op, and inside the step loop / a CUDA graph."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _ref_logits(obs, w1, b1, w2, b2, activation):
    """float64 reference of logit = w2 . act(W1 . nan_to_num(obs) + b1) + b2, and the magnitude the float16 input rounding scales with."""
    x = torch.nan_to_num(obs.double(), nan=0.0)
    pre = x @ w1.double().T + b1.double()
    hid = torch.relu(pre) if activation == "relu" else torch.tanh(pre)
    # sum of (|x w1| + |b1|) |w2|: what the 2^-11 relative roundings (inputs, weights, bias, hidden units, tanh) act on
    mag = (x.abs() @ w1.double().abs().T + b1.double().abs()) @ w2.double().abs()
    return hid @ w2.double() + b2, mag


@pytest.mark.parametrize("n,H,activation", [(1, 8, "relu"), (33, 16, "tanh"), (4099, 64, "relu"), (100_000, 32, "tanh")])
def test_policy_mlp_matches_a_torch_reference(golden, n, H, activation):
    from weather2alert_b200 import HeatAlertVectorEnv

    env = HeatAlertVectorEnv(n, weights=golden.weights("std"), store=golden.store(), seed=3)
    obs, _ = env.reset()
    for _ in range(3):  # a few steps, so that the agent slots are populated
        obs, *_ = env.step(torch.rand(n, device="cuda") < 0.5)
    assert bool(torch.isnan(obs[:, 27]).any())  # the significance slot: NaN where there is no alert
    g = torch.Generator(device="cuda").manual_seed(H)
    w1 = torch.randn(H, 29, generator=g, device="cuda") * 0.3
    w1[:, 17] *= 0.02  # day of summer runs to 152
    b1 = torch.randn(H, generator=g, device="cuda") * 0.1
    w2 = torch.randn(H, generator=g, device="cuda") * 0.5
    pol = env.mlp_policy(w1, b1, w2, -0.1, activation=activation)
    act = pol(obs).clone()
    ref, mag = _ref_logits(obs, w1, b1, w2, -0.1, activation)
    err = (pol.logits.double() - ref).abs()
    tol = 2.0 ** -9 * mag + 1e-5  # float16 inputs: both operands rounded to 11 significant bits, float32 accumulation
    assert bool((err <= tol).all()), float((err / tol).max())
    assert float(err.max()) < 0.05
    sure = ref.abs() > tol
    assert torch.equal(act[sure].bool(), ref[sure] > 0) and bool(sure.float().mean() > 0.9)
    # a float32 torch evaluation of the same policy agrees on those envs too
    ref32 = (torch.relu if activation == "relu" else torch.tanh)(torch.nan_to_num(obs) @ w1.T + b1) @ w2 - 0.1
    assert torch.equal(act[sure].bool(), ref32[sure] > 0)


def test_policy_mlp_stochastic_stream_and_errors(golden):
    from weather2alert_b200 import HeatAlertVectorEnv, _lib
    from weather2alert_b200.distributed import rollout_hash

    n, H = 20_000, 16
    env = HeatAlertVectorEnv(n, weights=golden.weights("std"), store=golden.store(), seed=4)
    obs, _ = env.reset()
    w1, b1, w2 = torch.zeros(H, 29, device="cuda"), torch.zeros(H, device="cuda"), torch.zeros(H, device="cuda")
    pol = env.mlp_policy(w1, b1, w2, -1.0, stochastic=True, seed=11)  # logit = -1 for every env: p = sigmoid(-1) = 0.269
    a0 = pol(obs).clone()
    a1 = pol(obs).clone()
    assert abs(float(a0.float().mean()) - 0.2689) < 0.01 and not torch.equal(a0, a1)  # the step counter advanced
    thr = np.uint64(np.float32(1.0) / (np.float32(1.0) + np.exp(np.float32(1.0))) * 4294967296.0)
    want = rollout_hash(11, np.arange(n), np.array([0]))[0] < thr
    assert (a0.cpu().numpy().astype(bool) != want).mean() < 1e-3  # __expf vs numpy in the last bit of the threshold
    with pytest.raises(_lib.W2AError, match="multiple of 8"):
        env.core.policy_mlp(obs, torch.zeros(12, 29, device="cuda"), torch.zeros(12, device="cuda"),
                            torch.zeros(12, device="cuda"), 0.0, a0)
    with pytest.raises(ValueError):
        env.mlp_policy(torch.zeros(H, 28), b1, w2, 0.0)


def test_mlp_policy_in_a_cuda_graph_equals_the_eager_loop():
    from weather2alert_b200 import HeatAlertVectorEnv, synthetic

    fips, zones = synthetic.synthetic_counties(20, seed=2)
    store = synthetic.make_store(fips, [2006, 2007], seed=3, zones=zones)
    weights = synthetic.make_weights(fips, 8, seed=4)
    n, H = 5000, 32
    g = torch.Generator(device="cuda").manual_seed(1)
    w1 = torch.randn(H, 29, generator=g, device="cuda") * 0.3
    w1[:, 17] *= 0.02
    b1, w2 = torch.randn(H, generator=g, device="cuda") * 0.1, torch.randn(H, generator=g, device="cuda") * 0.5
    kw = dict(weights=weights, store=store, seed=12, budget=6)
    eager = HeatAlertVectorEnv(n, autoreset=False, **kw)
    obs, _ = eager.reset()
    pol = eager.mlp_policy(w1, b1, w2, 0.05, activation="tanh")
    ret = torch.zeros(n, dtype=torch.float64, device="cuda")
    n_alerts = 0
    for _ in range(153):
        a = pol(obs)
        n_alerts += int(a.sum())
        obs, r, term, _, _ = eager.step(a)
        ret += r
    assert bool(term.all()) and 0 < n_alerts < 153 * n
    # captured: step + policy as ONE launch per iteration (w2a_step_mlp, the default), and as two launches
    for fuse in (True, False):
        graphed = HeatAlertVectorEnv(n, **kw)
        graphed.reset()
        gpol = graphed.mlp_policy(w1, b1, w2, 0.05, activation="tanh")
        gpol.fuse_step = fuse
        before = graphed.core.launch_count
        roll = graphed.graphed_rollout(gpol, reset_at_end=True)
        assert torch.equal(roll.replay(), ret), fuse
        assert torch.equal(gpol.logits, pol.logits) or fuse  # fused: one policy evaluation further (the terminal rows)
        del roll
    assert before > 0


@pytest.mark.parametrize("H,activation,stochastic", [(64, "relu", False), (16, "tanh", True), (40, "relu", True)])
def test_step_mlp_is_step_then_policy(golden, H, activation, stochastic):
    """`w2a_step_mlp` (one launch: the step of action k and the policy's action k + 1 from the fresh rows on chip) against
    `w2a_step` followed by `w2a_policy_mlp`, with the whole-summer tables (TABLE: the fused kernel) and without (COOP:
    the library's own two-launch form) -- observations, rewards, flags, next actions, logits and return sums identical."""
    from weather2alert_b200 import _lib
    from weather2alert_b200.core import HeatAlertCore

    store, weights = golden.store(), golden.weights("std")
    for tables in (True, False):
        core = HeatAlertCore(store, weights, "cuda:0", precompute=tables)
        assert core.has_tables == tables
        n = 20011 if tables else 3001
        rng = np.random.default_rng(5)
        col = rng.integers(0, len(weights.fips_list), n)  # every env on its own county's column: the tabulated path
        idx = {k: torch.from_numpy(np.asarray(v).astype(np.int32)).cuda() for k, v in dict(
            loc=[store.fips_index[weights.fips_list[c]] for c in col], year=rng.integers(0, store.shape[1], n),
            coef_col=col, coef_idx=rng.integers(0, weights.n_samples, n)).items()}
        g = torch.Generator(device="cuda").manual_seed(H)
        w1 = torch.randn(H, 29, generator=g, device="cuda") * 0.3
        w1[:, 17] *= 0.02
        b1, w2 = torch.randn(H, generator=g, device="cuda") * 0.1, torch.randn(H, generator=g, device="cuda") * 0.5
        kw = dict(activation=activation, stochastic=stochastic, seed=9)
        variant = _lib.STEP_TABLE if tables else _lib.STEP_COOP

        def buffers():
            st, ob = core.alloc_state(n), torch.zeros((n, 29), device="cuda")
            core.reset(st, ob, idx["loc"], idx["year"], idx["coef_col"], idx["coef_idx"])
            return dict(state=st, obs=ob, reward=torch.zeros(n, device="cuda"), term=torch.zeros(n, dtype=torch.uint8, device="cuda"),
                        rem=torch.zeros(n, dtype=torch.int32, device="cuda"), atb=torch.zeros(n, dtype=torch.uint8, device="cuda"),
                        a=[torch.zeros(n, dtype=torch.uint8, device="cuda") for _ in range(2)], logit=torch.zeros(n, device="cuda"),
                        ret=torch.zeros(n, dtype=torch.float64, device="cuda"))

        one, two = buffers(), buffers()
        for b in (one, two):
            core.policy_mlp(b["obs"], w1, b1, w2, -0.2, b["a"][0], logit=b["logit"], step=0, **kw)
        for k in range(12):
            cur, nxt = k & 1, (k + 1) & 1
            core.step_mlp(one["state"], one["a"][cur], one["obs"], one["reward"], one["term"], one["rem"], one["atb"], w1, b1, w2,
                          -0.2, one["a"][nxt], logit=one["logit"], ep_return=one["ret"], step=k + 1, variant=variant, **kw)
            core.step(two["state"], two["a"][cur], two["obs"], two["reward"], two["term"], two["rem"], two["atb"], variant=variant)
            two["ret"] += two["reward"]
            core.policy_mlp(two["obs"], w1, b1, w2, -0.2, two["a"][nxt], logit=two["logit"], step=k + 1, **kw)
            for key in ("state", "obs", "reward", "term", "rem", "atb", "logit", "ret"):
                assert torch.equal(one[key].view(torch.uint8), two[key].view(torch.uint8)), (tables, k, key)
            assert torch.equal(one["a"][nxt], two["a"][nxt]), (tables, k)
        assert 0 < int(one["a"][0].sum()) < n
    with pytest.raises(ValueError, match="alias"):
        core.step_mlp(one["state"], one["a"][0], one["obs"], one["reward"], one["term"], one["rem"], one["atb"], w1, b1, w2, 0.0,
                      one["a"][0])


def test_step_mlp_at_the_benchmarked_shape():
    """2^20 envs on the bench's own tables (746 x 11 x 153 x 100): five iterations of the one-launch form against
    step + policy as two launches -- every output byte identical -- and the step's part against the C oracle."""
    import bench
    from oracle import c_oracle
    from weather2alert_b200 import _lib
    from weather2alert_b200.core import HeatAlertCore

    store, weights = bench.make_tables()
    core = HeatAlertCore(store, weights, "cuda:0")
    n, H = 1 << 20, 64
    s = bench.index_streams(store, weights, n, 0, False, "interleaved")
    idx = {k: torch.from_numpy(v).cuda() for k, v in s.items()}
    g = torch.Generator(device="cuda").manual_seed(3)
    w1 = torch.randn(H, 29, generator=g, device="cuda") * 0.3
    w1[:, 17] *= 0.02
    b1, w2 = torch.randn(H, generator=g, device="cuda") * 0.1, torch.randn(H, generator=g, device="cuda") * 0.5

    def buffers():
        st, ob = core.alloc_state(n), torch.zeros((n, 29), device="cuda")
        core.reset(st, ob, idx["loc"], idx["year"], idx["coef_col"], idx["coef_idx"])
        return dict(state=st, obs=ob, reward=torch.zeros(n, device="cuda"), term=torch.zeros(n, dtype=torch.uint8, device="cuda"),
                    rem=torch.zeros(n, dtype=torch.int32, device="cuda"), atb=torch.zeros(n, dtype=torch.uint8, device="cuda"),
                    a=[torch.zeros(n, dtype=torch.uint8, device="cuda") for _ in range(2)], logit=torch.zeros(n, device="cuda"),
                    ret=torch.zeros(n, dtype=torch.float64, device="cuda"))

    one, two = buffers(), buffers()
    co = c_oracle.COracle(store.feat, store.sig_code, store.n_days, weights.coef, n)
    co.reset(s["loc"], s["year"], s["coef_col"], s["coef_idx"])
    for b in (one, two):
        core.policy_mlp(b["obs"], w1, b1, w2, 0.3, b["a"][0], logit=b["logit"])
    for k in range(5):
        cur, nxt = k & 1, (k + 1) & 1
        act = one["a"][cur].cpu().numpy()
        core.step_mlp(one["state"], one["a"][cur], one["obs"], one["reward"], one["term"], one["rem"], one["atb"], w1, b1, w2, 0.3,
                      one["a"][nxt], logit=one["logit"], ep_return=one["ret"], variant=_lib.STEP_TABLE)
        core.step(two["state"], two["a"][cur], two["obs"], two["reward"], two["term"], two["rem"], two["atb"], variant=_lib.STEP_TABLE)
        two["ret"] += two["reward"]
        core.policy_mlp(two["obs"], w1, b1, w2, 0.3, two["a"][nxt], logit=two["logit"])
        for key in ("state", "obs", "reward", "term", "rem", "atb", "logit", "ret"):
            assert torch.equal(one[key].view(torch.uint8), two[key].view(torch.uint8)), (k, key)
        assert torch.equal(one["a"][nxt], two["a"][nxt]), k
        o_ref, r_ref, d_ref, rem_ref, atb_ref = co.step(act)
        np.testing.assert_array_equal(one["obs"].cpu().numpy(), o_ref.astype(np.float32))
        np.testing.assert_allclose(one["reward"].cpu().numpy(), r_ref, rtol=1e-5, atol=1e-10)
        np.testing.assert_array_equal(one["rem"].cpu().numpy(), rem_ref)
    assert 0.01 < float(one["a"][0].float().mean()) < 0.99
