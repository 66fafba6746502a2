# small and mid-sized batches through the step path (CUDA graph per summer): env-steps/s and us/step, per variant
for v in thread coop; do for n in ${@:-1 32 256 1024 4096 16384 32768}; do
  python bench.py --workload cfg2_4096 --envs-per-gpu $n --variant $v --no-e2e --no-cpu-baseline 2>gpurun_out/small_err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', $n, '%.3e'%d['value'], '%.2f us/step'%(d['ms_per_step']*1e3))" || tail -2 gpurun_out/small_err.log
done; done
