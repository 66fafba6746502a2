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
    mag = (x.abs() @ w1.double().abs().T) @ w2.double().abs()  # sum of |x w1| |w2|: what a 2^-11 relative rounding acts on
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
    graphed = HeatAlertVectorEnv(n, **kw)
    graphed.reset()
    roll = graphed.graphed_rollout(graphed.mlp_policy(w1, b1, w2, 0.05, activation="tanh"), reset_at_end=True)
    assert torch.equal(roll.replay(), ret)
