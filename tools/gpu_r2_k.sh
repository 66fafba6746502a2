#!/bin/bash
# round 2, GPU session K: fused MLP policy kernel: tests + bench extra
set -x
O=gpurun_out/r2k
mkdir -p $O
( time python -m pytest tests/test_gpu_policy.py tests/test_gpu_vector.py -m gpu -q -x ) > $O/pytest.log 2>&1
tail -12 $O/pytest.log
python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-other-configs > $O/bench.json 2> $O/bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2k/bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['roofline']['frac'])
print(json.dumps(d['policy_in_loop'], indent=1))
PY
