# TABLE variant: parity, bench lines, launch list and one full capture of each of its two kernels
mkdir -p gpurun_out/table
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
for a in "--variant table" "--variant table --locality sorted" "--variant table --workload cfg5_8m --steps 612 --warmup 153" "--variant coop"; do
  python bench.py $a --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$a', '%.3e'%d['value'], 'ms/step %.4f'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'])"
done
CMD="python bench.py --variant table --no-graph --steps 12 --warmup 4 --no-e2e --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/table/plain.json 2>/dev/null && timeout 900 ncu --set full --clock-control none --import-source on -k regex:table_kernel -s 8 -c 2 -o gpurun_out/table/prof_table $CMD > gpurun_out/table/ncu_full.log 2>&1
ls -la gpurun_out/table
