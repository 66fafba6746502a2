#!/bin/bash
# round 2, GPU session C (1 GPU): full GPU test suite, L2-resident state A/B, memcheck attempt
set -x
O=gpurun_out/r2c
mkdir -p $O
( time python -m pytest tests -m gpu -x -q ) > $O/pytest.log 2>&1
tail -6 $O/pytest.log
Q="--no-extras --no-e2e --no-cpu-baseline --no-parity"
for rep in 1 2; do
  python bench.py $Q > $O/keep_on_$rep.json 2>> $O/var.err
  W2A_KEEP_STATE_MAX=0 python bench.py $Q > $O/keep_off_$rep.json 2>> $O/var.err
done
python bench.py $Q --locality sorted > $O/keep_on_sorted.json 2>> $O/var.err
W2A_KEEP_STATE_MAX=0 python bench.py $Q --locality sorted > $O/keep_off_sorted.json 2>> $O/var.err
python bench.py $Q --envs-per-gpu 2097152 > $O/keep_on_2m.json 2>> $O/var.err
W2A_KEEP_STATE_MAX=0 python bench.py $Q --envs-per-gpu 2097152 > $O/keep_off_2m.json 2>> $O/var.err
python bench.py --steps 20 --warmup 5 > $O/bench_driver.json 2> $O/bench_driver.err
python tools/gpu_memcheck.py > $O/memcheck_plain.log 2>&1 && timeout 600 compute-sanitizer --tool memcheck python tools/gpu_memcheck.py > $O/memcheck.log 2>&1
tail -5 $O/memcheck.log
python - <<'PY'
import json, glob, os
for f in sorted(glob.glob('gpurun_out/r2c/*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(os.path.basename(f), d.get('value'), d.get('ms_per_step'), d.get('roofline', {}).get('frac'))
    except Exception as e:
        print(os.path.basename(f), 'ERR', e)
PY
