timeout 900 python -m pytest tests/test_gpu_rollout.py -x -q > gpurun_out/pytest_rollout.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_rollout.log
tail -25 gpurun_out/pytest_rollout.log
python - <<'PY' > gpurun_out/rollout_bench.txt 2>&1
import sys, time, torch, numpy as np
sys.path.insert(0, '.')
import bench
from weather2alert_b200.core import HeatAlertCore
store, weights = bench.make_tables()
core = HeatAlertCore(store, weights, 'cuda:0')
for n, loc in ((1<<20,'interleaved'), (1<<20,'sorted'), (1<<22,'interleaved'), (1<<22,'sorted')):
    s = bench.index_streams(store, weights, n, 0, False, loc)
    idx = {k: torch.from_numpy(v).cuda() for k, v in s.items()}
    state = core.alloc_state(n); obs = torch.zeros((n,29), device='cuda')
    ret = torch.zeros(n, dtype=torch.float64, device='cuda'); stats = torch.zeros(15, dtype=torch.float64, device='cuda')
    ts = []
    for it in range(4):
        core.reset(state, obs, idx['loc'], idx['year'], idx['coef_col'], idx['coef_idx'])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); core.rollout(state, 153, ret, seed=it, p_alert=0.5, stats=stats); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = min(ts[1:])
    print(f"rollout n={n} {loc}: {ms:.3f} ms per 153-step rollout -> {n*153/ms*1e3:.3e} env-steps/s  (all: {[round(t,3) for t in ts]})", flush=True)
PY
cat gpurun_out/rollout_bench.txt
