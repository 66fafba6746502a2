#!/bin/bash
# round 2, GPU session F: full tests on the sector-table kernel; load-form variants
set -x
O=gpurun_out/r2f
mkdir -p $O
( time python -m pytest tests -m gpu -q ) > $O/pytest.log 2>&1
tail -8 $O/pytest.log
Q="--no-extras --no-e2e --no-cpu-baseline --no-parity"
for v in base z256 znoalloc rownoalloc allnoalloc; do
  W2A_LIB_PATH=build/variants/libw2a_$v.so python bench.py $Q > $O/var_$v.json 2>> $O/var.err
  W2A_LIB_PATH=build/variants/libw2a_$v.so python bench.py $Q --locality sorted > $O/sorted_$v.json 2>> $O/var.err
done
python - <<'PY'
import json, glob, os
for f in sorted(glob.glob('gpurun_out/r2f/*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(os.path.basename(f).ljust(26), round(d.get('ms_per_step')*1000,2), round(d.get('roofline', {}).get('frac'),3))
    except Exception as e:
        print(os.path.basename(f), 'ERR', e)
PY
