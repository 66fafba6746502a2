"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck): ragged sizes, masked
reset, all three step variants (TABLE on the precomputed whole-summer tables), host entry points, rollout with bank,
hashed actions and the in-launch policy, and the feature-engineering kernels."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from weather2alert_b200 import synthetic
from weather2alert_b200.core import HeatAlertCore

fips, zones = synthetic.synthetic_counties(12, seed=1)
nd = np.full((12, 2), 153, np.uint8); nd[3, 1] = 5; nd[7, 0] = 1
store = synthetic.make_store(fips, [2006, 2007], seed=2, zones=zones, n_days=nd)
weights = synthetic.make_weights(fips, 7, seed=3)
core = HeatAlertCore(store, weights, "cuda:0")
rng = np.random.default_rng(0)
for n in (1, 31, 33, 257, 2049, 4100):
    idx = {k: torch.from_numpy(v.astype(np.int32)).cuda() for k, v in dict(
        loc=rng.integers(0, 12, n), year=rng.integers(0, 2, n), col=rng.integers(0, 12, n), p=rng.integers(0, 7, n)).items()}
    for variant in (1, 2, 3):
        state = core.alloc_state(n); obs = torch.zeros((n, 29), device="cuda")
        rew = torch.zeros(n, device="cuda"); term = torch.zeros(n, dtype=torch.uint8, device="cuda")
        rem = torch.zeros(n, dtype=torch.int32, device="cuda"); atb = torch.zeros(n, dtype=torch.uint8, device="cuda")
        core.reset(state, obs, idx["loc"], idx["year"], idx["col"], idx["p"], remaining=rem)
        for k in range(12):
            a = torch.from_numpy((rng.random(n) < 0.5).astype(np.uint8)).cuda()
            core.step(state, a, obs, rew, term, rem, atb, variant=variant)
        mask = torch.from_numpy((rng.random(n) < 0.5).astype(np.uint8)).cuda()
        core.reset(state, obs, idx["loc"], idx["year"], idx["col"], idx["p"], mask=mask)
        ret = torch.zeros(n, dtype=torch.float64, device="cuda"); stats = torch.zeros(15, dtype=torch.float64, device="cuda")
        acts = torch.from_numpy((rng.random((20, n)) < 0.5).astype(np.uint8)).cuda()
        core.rollout(state, 20, ret, actions=acts, stats=stats)
        core.rollout(state, 200, ret, seed=5, p_alert=0.4, stats=stats)
        core.reset(state, obs, idx["loc"], idx["year"], idx["col"], idx["p"])
        core.rollout_policy(state, 153, ret, torch.linspace(-1, 1, 30, dtype=torch.float64).cuda(), stochastic=True, seed=3, stats=stats)
    core.host_envs_create(n)
    h = {k: v.cpu().numpy() for k, v in idx.items()}
    ob = np.zeros((n, 29), np.float32); rw = np.zeros(n, np.float32); tm = np.zeros(n, np.uint8)
    core.reset_host(h["loc"], h["year"], h["col"], h["p"], ob)
    for k in range(3):
        core.step_host((rng.random(n) < 0.5).astype(np.uint8), ob, rw, tm)
from weather2alert_b200.features import FeatureEngineer, holiday_flags
days = np.concatenate([np.arange(13270, 13270 + m) for m in (153, 40, 1, 153 * 3)]).astype(np.int32)
start = np.array([0, 153, 193, 194, 194 + 459], np.int64)
m = len(days)
FeatureEngineer("cuda:0").engineer(start, days, rng.normal(28, 4, m), rng.random(m) < 0.1, np.full(m, np.nan), holiday_flags(days))
torch.cuda.synchronize()
print("memcheck workload done; launches", core.launch_count)
