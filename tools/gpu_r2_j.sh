#!/bin/bash
# round 2, GPU session J: table-driven rollout kernels: tests + bench (rollout extras)
set -x
O=gpurun_out/r2j
mkdir -p $O
( time python -m pytest tests -m gpu -q ) > $O/pytest.log 2>&1
tail -6 $O/pytest.log
python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > $O/bench.json 2> $O/bench.err
python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --locality sorted --no-other-configs > $O/bench_sorted.json 2>> $O/bench.err
python - <<'PY'
import json
for f in ('bench', 'bench_sorted'):
    d = json.loads(open(f'gpurun_out/r2j/{f}.json').read().strip().splitlines()[-1])
    print(f, d['ms_per_step'], d['roofline']['frac'])
    print(json.dumps(d['rollout'])[:900])
    print(json.dumps(d.get('vector_env_default'))[:300])
    print(json.dumps(d['other_configs'])[:1500])
PY
