"""One rollout launch at 2^20 envs for ncu (interleaved and sorted)."""
import sys, torch
sys.path.insert(0, '.')
import bench
from weather2alert_b200.core import HeatAlertCore
store, weights = bench.make_tables()
core = HeatAlertCore(store, weights, 'cuda:0')
n = 1 << 20
for loc in ('interleaved', 'sorted'):
    s = bench.index_streams(store, weights, n, 0, False, loc)
    idx = {k: torch.from_numpy(v).cuda() for k, v in s.items()}
    state = core.alloc_state(n); obs = torch.zeros((n, 29), device='cuda')
    ret = torch.zeros(n, dtype=torch.float64, device='cuda'); stats = torch.zeros(15, dtype=torch.float64, device='cuda')
    for it in range(2):
        core.reset(state, obs, idx['loc'], idx['year'], idx['coef_col'], idx['coef_idx'])
        core.rollout(state, 153, ret, seed=it, p_alert=0.5, stats=stats)
    torch.cuda.synchronize()
print("done")
