#!/bin/bash
# round 2, GPU session D: what the record gather costs -- request count vs HBM misses
set -x
O=gpurun_out/r2d
mkdir -p $O
Q="--no-extras --no-e2e --no-cpu-baseline --no-parity"
for v in base k16 k32 k48 k2 k4 k6; do
  W2A_LIB_PATH=build/variants/libw2a_$v.so python bench.py $Q > $O/var_$v.json 2>> $O/var.err
done
for v in base k16 k32 k2; do
  W2A_LIB_PATH=build/variants/libw2a_$v.so python bench.py $Q --locality sorted > $O/sorted_$v.json 2>> $O/var.err
done
python - <<'PY'
import json, glob, os
for f in sorted(glob.glob('gpurun_out/r2d/*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(os.path.basename(f), d.get('value'), d.get('ms_per_step'), d.get('roofline', {}).get('frac'))
    except Exception as e:
        print(os.path.basename(f), 'ERR', e)
PY
