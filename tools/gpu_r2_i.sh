#!/bin/bash
# round 2, GPU session I: cooperative sector gather (8 lanes x 4 B per env, 4 envs per instruction)
set -x
O=gpurun_out/r2i
mkdir -p $O
( W2A_LIB_PATH=build/variants/libw2a_zcoop.so python -m pytest tests/test_gpu_parity.py tests/test_gpu_vector.py -m gpu -q -x ) > $O/pytest_zcoop.log 2>&1
tail -4 $O/pytest_zcoop.log
Q="--no-extras --no-e2e --no-cpu-baseline --no-parity"
for v in base zcoop zcoop_rows4; do
  W2A_LIB_PATH=build/variants/libw2a_$v.so python bench.py $Q > $O/var_$v.json 2>> $O/var.err
  W2A_LIB_PATH=build/variants/libw2a_$v.so python bench.py $Q --locality sorted > $O/sorted_$v.json 2>> $O/var.err
  W2A_LIB_PATH=build/variants/libw2a_$v.so python bench.py $Q --workload cfg5_8m > $O/cfg5_$v.json 2>> $O/var.err
done
python - <<'PY'
import json, glob, os
for f in sorted(glob.glob('gpurun_out/r2i/*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(os.path.basename(f).ljust(30), round(d.get('ms_per_step')*1000,2), round(d.get('roofline', {}).get('frac'),3))
    except Exception as e:
        print(os.path.basename(f), 'ERR', e)
PY
