"""The MLP policy kernel alone at 2^20 envs: CUDA-event time over alternating observation buffers (2 x 122 MB, so that
no launch finds its input in L2), relu and tanh, H = 16 / 32 / 64.  Also the target of
`ncu --set full -k regex:policy_mlp_kernel -c 1` (profiles/r02/policy_mlp_*)."""
import json
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from weather2alert_b200.core import HeatAlertCore  # noqa: E402

store, weights = bench.make_tables(years=2, samples=2)
core = HeatAlertCore(store, weights, "cuda:0", precompute=False)
n = 1 << 20
obs = [torch.randn((n, 29), device="cuda") for _ in range(2)]
for o in obs:
    o[:, 27] = float("nan")
act = torch.zeros(n, dtype=torch.uint8, device="cuda")
out = {}
combos = [(64, "relu")] if "--one" in sys.argv else [(h, a) for a in ("relu", "tanh") for h in (16, 32, 64)]
for H, a in combos:
    w1, b1, w2 = torch.randn(H, 29, device="cuda") * 0.3, torch.randn(H, device="cuda") * 0.1, torch.randn(H, device="cuda")
    for i in range(4):
        core.policy_mlp(obs[i & 1], w1, b1, w2, 0.0, act, activation=a)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(20):
        core.policy_mlp(obs[i & 1], w1, b1, w2, 0.0, act, activation=a)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    out[f"{a}_H{H}"] = {"us": round(us, 2), "GB/s": round(n * 117 / us / 1e3, 1)}
print(json.dumps(out))
