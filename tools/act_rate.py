import json, statistics, sys
sys.path.insert(0, "/root/repo")
import torch, bench
from weather2alert_b200 import _lib
from weather2alert_b200.core import HeatAlertCore
store, weights = bench.make_tables()
dev = torch.device("cuda:0")
core = HeatAlertCore(store, weights, dev)
T = store.shape[2]; n = 1 << 20
idx = {k: torch.from_numpy(v).to(dev) for k, v in bench.index_streams(store, weights, n, 0, False, "interleaved").items()}
out = {}
for p in (0.0, 0.05, 0.5, 1.0):
    actions = (torch.rand((T, n), device=dev) < p).to(torch.uint8)
    es = bench.EnvSet(core, idx, n, T, actions, _lib.STEP_TABLE)
    es.reset(); es.run(3)
    tm, _, _, _ = bench.timed_steps(es, T, T, 1, None, target_ms=30.0)
    out[str(p)] = round(statistics.median(tm) / T * 1e3, 2)
print(json.dumps(out))
