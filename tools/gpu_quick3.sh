python -m pytest tests/test_gpu_parity.py tests/test_gpu_vector.py -m gpu -x -q 2>&1 | tail -3
for a in "--variant table" "--variant table --locality sorted" "--variant table --workload cfg5_8m --steps 612 --warmup 153"; do
  python bench.py $a --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$a', '%.3e'%d['value'], 'ms/step %.4f'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'])"
done
bash tools/gpu_warm_launches.sh --variant table | tail -4
