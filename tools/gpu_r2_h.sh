#!/bin/bash
# round 2, GPU session H: evidence run on the final kernels (tests, bench lines, knock-outs, ncu launch list + full capture)
set -x
O=gpurun_out/r2h
mkdir -p $O
( time python -m pytest tests -m gpu -q ) > $O/pytest.log 2>&1
tail -4 $O/pytest.log
python bench.py --steps 20 --warmup 5 > $O/bench_driver_flags.json 2> $O/bench_driver_flags.err
python bench.py > $O/bench_default.json 2> $O/bench_default.err
python bench.py --steps 20 --warmup 5 --weights-layout nn --no-other-configs > $O/bench_nn_layout.json 2> $O/bench_nn_layout.err
Q="--no-extras --no-e2e --no-cpu-baseline --no-parity"
for w in cfg2_4096 cfg4_2m cfg5_8m; do
  python bench.py $Q --workload $w > $O/bench_$w.json 2>> $O/var.err
done
python bench.py $Q --workload cfg5_8m --locality sorted > $O/bench_cfg5_8m_sorted.json 2>> $O/var.err
for v in base k1 k2 k4 k6 k8 k9 k15 k16 k64 k128 rows4 z128; do
  W2A_LIB_PATH=build/variants/libw2a_$v.so python bench.py $Q > $O/var_$v.json 2>> $O/var.err
done
W2A_PDL=0 python bench.py $Q > $O/var_nopdl.json 2>> $O/var.err
W2A_KEEP_STATE_MAX=0 python bench.py $Q > $O/var_nokeep.json 2>> $O/var.err
python bench.py $Q --no-graph > $O/var_nograph.json 2>> $O/var.err
python bench.py $Q --variant coop > $O/var_coop.json 2>> $O/var.err
CMD="python bench.py $Q --steps 20 --warmup 5 --no-graph"
$CMD > $O/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv $CMD > $O/ncu_launches.log 2>&1
$CMD > $O/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:step_table_kernel -s 12 -c 2 -o $O/step_table $CMD > $O/ncu_full.log 2>&1
python - <<'PY'
import json, glob, os
for f in sorted(glob.glob('gpurun_out/r2h/*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(os.path.basename(f).ljust(30), round(d.get('ms_per_step')*1000,2), round(d.get('roofline', {}).get('frac'),3))
    except Exception as e:
        print(os.path.basename(f), 'ERR', e)
PY
