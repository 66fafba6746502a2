timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log; tail -4 gpurun_out/pytest.log
for args in "" "--locality sorted" "--workload cfg5_8m --steps 612" "--workload cfg5_8m --steps 612 --locality sorted" "--workload cfg4_2m" "--workload cfg4_2m --locality sorted"; do
  python bench.py --no-e2e --no-cpu-baseline $args 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$args |', '%.3e'%d['value'], 'ms/step %.4f'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'], 'rollout %.3e'%d['rollout']['value'], d['rollout']['same_returns_as_stepping'])"
done
