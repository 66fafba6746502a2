"""Step time of small env sets on the bench's real-shape tables (746 x 11 x 153, 100 samples): one CUDA graph per
summer, device-resident, TABLE variant.  `W2A_SMALL_TABLE=<groups per SM>` moves the switch to the one-warp-per-CTA
instance (0 = never).  Usage: python tools/small_sets.py [n ...]"""
import json
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from weather2alert_b200 import _lib  # noqa: E402
from weather2alert_b200.core import HeatAlertCore  # noqa: E402

sizes = [int(a) for a in sys.argv[1:]] or [32, 1024, 4096, 8192, 16384, 32768, 65536]
store, weights = bench.make_tables()
dev = torch.device("cuda:0")
core = HeatAlertCore(store, weights, dev)
T = store.shape[2]
out = {"W2A_SMALL_TABLE": os.environ.get("W2A_SMALL_TABLE", "default")}
for n in sizes:
    idx = {k: torch.from_numpy(v).to(dev) for k, v in bench.index_streams(store, weights, n, 0, False, "interleaved").items()}
    actions = (torch.rand((T, n), device=dev) < 0.3).to(torch.uint8)
    es = bench.EnvSet(core, idx, n, T, actions, _lib.STEP_TABLE)
    es.reset()
    es.run(3)
    tm, _, _, _ = bench.timed_steps(es, T, T, 1, None, target_ms=20.0)
    out[str(n)] = round(statistics.median(tm) / T * 1e3, 3)  # microseconds per step
print(json.dumps(out))
