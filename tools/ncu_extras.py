"""One table-form rollout, one policy-rollout and one MLP-policy launch at 2^20 envs on the bench's tables, for
`ncu --set full -k regex:'rollout_kernel|policy_mlp_kernel'` (profiles/r02/extras_*)."""
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from weather2alert_b200.core import HeatAlertCore  # noqa: E402

store, weights = bench.make_tables()
core = HeatAlertCore(store, weights, "cuda:0")
n = 1 << 20
s = bench.index_streams(store, weights, n, 0, False)
idx = {k: torch.from_numpy(v).cuda() for k, v in s.items()}
state, obs = core.alloc_state(n), torch.zeros((n, 29), device="cuda")
ret, stats = torch.zeros(n, dtype=torch.float64, device="cuda"), torch.zeros(15, dtype=torch.float64, device="cuda")
acts = (torch.rand((153, n), device="cuda") < 0.5).to(torch.uint8)
theta = torch.zeros(30, dtype=torch.float64)
theta[0], theta[17], theta[25], theta[28], theta[29] = 2.0, 0.004, 0.05, -0.3, -2.2
theta = theta.cuda()
for _ in range(2):
    core.reset(state, obs, idx["loc"], idx["year"], idx["coef_col"], idx["coef_idx"])
    core.rollout(state, 153, ret, actions=acts, stats=stats)
    core.reset(state, obs, idx["loc"], idx["year"], idx["coef_col"], idx["coef_idx"])
    core.rollout_policy(state, 153, ret, theta, stochastic=True, seed=1, stats=stats)
    w1, b1, w2 = torch.randn(64, 29, device="cuda") * 0.3, torch.zeros(64, device="cuda"), torch.randn(64, device="cuda")
    core.policy_mlp(obs, w1, b1, w2, 0.0, torch.zeros(n, dtype=torch.uint8, device="cuda"))
torch.cuda.synchronize()
print("done", float(stats[0]))
