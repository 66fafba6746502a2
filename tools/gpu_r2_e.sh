#!/bin/bash
# round 2, GPU session E: one-sector float32 logit table (+ float64 fallback): tests, bench, variants
set -x
O=gpurun_out/r2e
mkdir -p $O
( time python -m pytest tests -m gpu -x -q ) > $O/pytest.log 2>&1
tail -6 $O/pytest.log
python bench.py --steps 20 --warmup 5 > $O/bench_driver.json 2> $O/bench_driver.err
Q="--no-extras --no-e2e --no-cpu-baseline --no-parity"
for v in base k16 k2 k4 zkeep mb6; do
  W2A_LIB_PATH=build/variants/libw2a_$v.so python bench.py $Q > $O/var_$v.json 2>> $O/var.err
done
python bench.py $Q --locality sorted > $O/sorted.json 2>> $O/var.err
python bench.py $Q --workload cfg5_8m > $O/cfg5.json 2>> $O/var.err
python bench.py $Q --weights-layout nn > $O/nn.json 2>> $O/var.err
python - <<'PY'
import json, glob, os
for f in sorted(glob.glob('gpurun_out/r2e/*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(os.path.basename(f).ljust(22), round(d.get('ms_per_step')*1000,2), round(d.get('roofline', {}).get('frac'),3))
    except Exception as e:
        print(os.path.basename(f), 'ERR', e)
PY
