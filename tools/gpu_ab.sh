# usage: gpu_ab.sh libA.so libB.so ...   A/B the TABLE variant between prebuilt libraries (bench + warm per-kernel times)
for lib in "$@"; do
  echo "== $lib"
  for a in "--variant table" "--variant table --locality sorted" "--variant table --workload cfg5_8m --steps 612 --warmup 153"; do
    W2A_LIB_PATH=$lib python bench.py $a --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$a', '%.3e'%d['value'], 'ms/step %.4f'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'])"
  done
done
