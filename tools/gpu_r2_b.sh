#!/bin/bash
# round 2, GPU session B: bulk-prefetch variant vs base; ncu launch list + full capture of step_table_kernel
set -x
O=gpurun_out/r2b
mkdir -p $O
Q="--no-extras --no-e2e --no-cpu-baseline --no-parity"
for v in base pf3 base pf3; do
  W2A_LIB_PATH=build/variants/libw2a_$v.so python bench.py $Q > $O/var_${v}_$RANDOM.json 2>> $O/var.err
done
W2A_LIB_PATH=build/variants/libw2a_pf3.so python bench.py $Q --locality sorted > $O/var_pf3_sorted.json 2>> $O/var.err
W2A_LIB_PATH=build/variants/libw2a_pf3.so python bench.py $Q --workload cfg5_8m > $O/var_pf3_cfg5.json 2>> $O/var.err
CMD="python bench.py $Q --steps 20 --warmup 5 --no-graph"
$CMD > $O/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv $CMD > $O/ncu_launches.log 2>&1
$CMD > $O/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:step_table_kernel -s 12 -c 2 -o $O/step_table $CMD > $O/ncu_full.log 2>&1
tail -3 $O/ncu_full.log
python - <<'PY'
import json, glob, os
for f in sorted(glob.glob('gpurun_out/r2b/*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(os.path.basename(f), d.get('value'), d.get('ms_per_step'), d.get('roofline', {}).get('frac'))
    except Exception as e:
        print(os.path.basename(f), 'ERR', e)
PY
