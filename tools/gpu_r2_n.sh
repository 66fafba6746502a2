#!/bin/bash
# round 2, GPU session N: the evidence of the final tree (both bench arms as the driver runs them, default bench, launch list)
set -x
O=gpurun_out/r2n
mkdir -p $O
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/bench_reference.json 2> $O/bench_reference.err
python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_final.json 2> $O/bench_final.err
python bench.py > $O/bench_default.json 2> $O/bench_default.err
python bench.py --steps 20 --warmup 5 --weights-layout nn --no-other-configs > $O/bench_nn_layout.json 2> $O/bench_nn_layout.err
python bench.py --no-extras --no-e2e --no-cpu-baseline --no-parity --workload cfg2_4096 > $O/bench_cfg2_4096.json 2> $O/var.err
CMD="python bench.py --no-extras --no-e2e --no-cpu-baseline --no-parity --steps 20 --warmup 5 --no-graph"
$CMD > $O/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv $CMD > $O/ncu_launches.log 2>&1
python tools/launch_summary.py $O/launches.csv "$CMD" > $O/launches_summary.txt
cat $O/launches_summary.txt | head -8
python - <<'PY'
import json, glob, os
for f in sorted(glob.glob('gpurun_out/r2n/*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(os.path.basename(f).ljust(28), d.get('value'), d.get('ms_per_step'), (d.get('roofline') or {}).get('frac'), (d.get('e2e') or {}).get('value'))
    except Exception as e:
        print(os.path.basename(f), 'ERR', e)
PY
