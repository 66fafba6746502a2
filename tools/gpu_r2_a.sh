#!/bin/bash
# round 2, GPU session A: tests, driver-flag bench, default bench, tuning variants of the table step kernel
set -x
O=gpurun_out/r2a
mkdir -p $O
nvidia-smi -L > $O/gpu.txt
( time python -m pytest tests -m gpu -x -q ) > $O/pytest.log 2>&1
tail -5 $O/pytest.log
python bench.py --steps 20 --warmup 5 > $O/bench_driver.json 2> $O/bench_driver.err
tail -3 $O/bench_driver.err
Q="--no-extras --no-e2e --no-cpu-baseline --no-parity"
for v in base pf1 pf2 zkeep mb6 rows4mb6 rows1 w8 w2 k1 k2 k4 k8 k6 k9 k15; do
  W2A_LIB_PATH=build/variants/libw2a_$v.so python bench.py $Q > $O/var_$v.json 2> $O/var_$v.err
done
W2A_PDL=0 python bench.py $Q > $O/var_nopdl.json 2> $O/var_nopdl.err
W2A_PDL=0 python bench.py $Q --steps 20 --warmup 5 > $O/var_nopdl_k20.json 2> $O/var_nopdl_k20.err
python bench.py $Q --no-graph > $O/var_nograph.json 2> $O/var_nograph.err
python bench.py $Q --locality sorted > $O/var_sorted.json 2> $O/var_sorted.err
python bench.py $Q --workload cfg5_8m > $O/var_cfg5.json 2> $O/var_cfg5.err
python bench.py $Q --variant coop > $O/var_coop.json 2> $O/var_coop.err
python - <<'PY'
import json, glob, os
for f in sorted(glob.glob('gpurun_out/r2a/*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(os.path.basename(f), d.get('value'), d.get('ms_per_step'), d.get('roofline', {}).get('frac'), d.get('timing', {}).get('repeats'))
    except Exception as e:
        print(os.path.basename(f), 'ERR', e)
PY
