"""`a = mlp_policy(obs); env.step(a)` at 2^20 envs on the bench's tables: the graphed loop timed with CUDA events, then a
few eager iterations for `ncu --metrics gpu__time_duration.sum --cache-control none` (which kernel costs what IN the loop)."""
import json
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from weather2alert_b200 import HeatAlertVectorEnv  # noqa: E402

store, weights = bench.make_tables()
n = 1 << 20
venv = HeatAlertVectorEnv(n, weights=weights, store=store, seed=5, device="cuda:0", sort_episodes="--sorted" in sys.argv)
venv.reset()
g = torch.Generator(device="cuda").manual_seed(int(next((a[6:] for a in sys.argv if a.startswith("--seed")), 1)))
B2 = float(next((a[4:] for a in sys.argv if a.startswith("--b2")), 0.0))
H = int(next((a[2:] for a in sys.argv if a.startswith("-H")), 64))
w1 = torch.randn(H, 29, generator=g, device="cuda") * 0.3
w2 = torch.randn(H, generator=g, device="cuda")
pol = venv.mlp_policy(w1, torch.zeros(H, device="cuda"), w2, B2)
pol.fuse_step = "--two-launches" not in sys.argv
if "--fused-eager" in sys.argv:  # the one-launch form without a graph, for `ncu -k regex:step_mlp_kernel`
    cur, nxt = pol(venv._obs), torch.zeros(n, dtype=torch.uint8, device="cuda")
    ret = torch.zeros(n, dtype=torch.float64, device="cuda")
    for k in range(8):
        venv.core.step_mlp(venv._state, cur, venv._obs, venv._reward, venv._term.view(torch.uint8), venv._rem,
                           venv._atb.view(torch.uint8), next_action=nxt, logit=pol.logits, ep_return=ret, step=k + 1,
                           variant=venv._variant(), **pol.mlp)
        cur, nxt = nxt, cur
    torch.cuda.synchronize()
    print("fused eager done", float(ret.sum()))
elif "--eager" in sys.argv:
    obs = venv._obs
    for _ in range(6):
        a = pol(obs)
        obs, *_ = venv.step(a)
    torch.cuda.synchronize()
    print("eager done")
else:
    roll = venv.graphed_rollout(pol, reset_at_end=True)
    roll.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(3):
        roll.replay()
    e1.record()
    torch.cuda.synchronize()
    print(json.dumps({"us_per_step": round(e0.elapsed_time(e1) / 3 / roll.steps * 1e3, 2), "alert_rate_last_step": round(float(pol.out.float().mean()), 3), "mean_return": round(float(roll.returns.mean()), 1)}))
