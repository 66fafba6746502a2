#!/bin/bash
# round 2, GPU session L: what the driver runs at round end, on the final tree: GPU tests, smoke, reference arm, bench
set -x
O=gpurun_out/r2l
mkdir -p $O
( time python -m pytest tests -x -q -m gpu ) > $O/pytest.log 2>&1
tail -4 $O/pytest.log
( time python -c "import __graft_entry__ as g; g.smoke()" ) > $O/smoke.log 2>&1
tail -6 $O/smoke.log
( time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > $O/bench_reference.json 2> $O/bench_reference.err
( time python bench.py --gpus 1 --steps 20 --warmup 5 ) > $O/bench_final.json 2> $O/bench_final.err
tail -4 $O/bench_final.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2l/bench_final.json').read().strip().splitlines()[-1])
r = json.loads(open('gpurun_out/r2l/bench_reference.json').read().strip().splitlines()[-1])
print('ours', d['value'], d['ms_per_step'], d['roofline']['frac'], 'e2e', d['e2e']['value'], d['e2e']['frac_of_link'])
print('reference', r['value'], 'ratio e2e', d['e2e']['value'] / r['value'], 'ratio value', d['value'] / r['value'])
print('cpu_baseline', d['cpu_baseline'])
print('sorted', d['sorted']['roofline_frac'], 'venv', d['vector_env_default']['roofline_frac'])
print('rollout', d['rollout']['value'], d['rollout']['policy_deterministic']['value'], d['rollout']['policy_stochastic']['value'])
print('policy loop', d['policy_in_loop']['value'], d['policy_in_loop']['fused_mlp_policy']['value'])
PY
