"""Where a whole summer's time goes at a given env-set size: the reset alone and every one of the 153 steps, each
bracketed by CUDA events (host launches, interleaved slots, TABLE).  usage: python tools/summer_profile.py [log2 n]"""
import json
import sys

sys.path.insert(0, ".")
import torch  # noqa: E402

import bench  # noqa: E402
from weather2alert_b200 import _lib  # noqa: E402
from weather2alert_b200.core import HeatAlertCore  # noqa: E402

n = 1 << int(sys.argv[1] if len(sys.argv) > 1 else 23)
store, weights = bench.make_tables()
dev = torch.device("cuda:0")
core = HeatAlertCore(store, weights, dev)
T = store.shape[2]
idx = {k: torch.from_numpy(v).to(dev) for k, v in bench.index_streams(store, weights, n, 0, False, "interleaved").items()}
actions = (torch.rand((T, n), device=dev) < 0.5).to(torch.uint8)
es = bench.EnvSet(core, idx, n, T, actions, _lib.STEP_TABLE, use_graph=False)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(T + 2)]
times = None
for rep in range(3):
    torch.cuda.synchronize()
    ev[0].record()
    es.reset()
    ev[1].record()
    for t in range(T):
        es.step(actions[t])
        ev[t + 2].record()
    torch.cuda.synchronize()
    times = [ev[i].elapsed_time(ev[i + 1]) * 1e3 for i in range(T + 1)]
steps = times[1:]
print(json.dumps({"n": n, "reset_us": round(times[0], 1), "step_us_first5": [round(x, 1) for x in steps[:5]],
                  "step_us_last5": [round(x, 1) for x in steps[-5:]], "median": round(sorted(steps)[T // 2], 1),
                  "mean": round(sum(steps) / T, 1), "max": round(max(steps), 1), "argmax": steps.index(max(steps)),
                  "by_30_days": [round(sum(steps[i:i + 30]) / len(steps[i:i + 30]), 1) for i in range(0, T, 30)]}))
