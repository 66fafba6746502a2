# final evidence for profiles/: launch list + full capture of the two TABLE-variant kernels (the default at 2^20 envs)
# and of the COOP kernel, at the default bench command
mkdir -p gpurun_out/final
CMD="python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/final/plain.json 2> gpurun_out/final/plain.err && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/final/launches.csv $CMD > gpurun_out/final/ncu_launches.log 2>&1
timeout 300 $CMD > gpurun_out/final/plain2.json 2> gpurun_out/final/plain2.err && timeout 900 ncu --set full --clock-control none --import-source on -k regex:table_kernel -s 16 -c 2 -o gpurun_out/final/prof_step_table $CMD > gpurun_out/final/ncu_full.log 2>&1
timeout 300 $CMD --variant coop > gpurun_out/final/plain3.json 2> gpurun_out/final/plain3.err && timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_coop -s 8 -c 1 -o gpurun_out/final/prof_step_coop $CMD --variant coop > gpurun_out/final/ncu_full_coop.log 2>&1
timeout 600 python bench.py > gpurun_out/final/bench_default.json 2> gpurun_out/final/bench_default.err
timeout 300 python bench.py --workload cfg2_4096 --no-cpu-baseline > gpurun_out/final/bench_cfg2.json 2>/dev/null
timeout 300 python bench.py --workload cfg4_2m --no-cpu-baseline > gpurun_out/final/bench_cfg4.json 2>/dev/null
timeout 300 python bench.py --workload cfg5_8m --no-cpu-baseline --no-e2e --steps 612 --warmup 153 > gpurun_out/final/bench_cfg5.json 2>/dev/null
timeout 300 python bench.py --variant coop --no-cpu-baseline --no-e2e > gpurun_out/final/bench_coop.json 2>/dev/null
timeout 300 python bench.py --locality sorted --no-cpu-baseline --no-e2e > gpurun_out/final/bench_sorted.json 2>/dev/null
timeout 300 python bench.py --workload cfg5_8m --locality sorted --no-cpu-baseline --no-e2e --steps 612 --warmup 153 > gpurun_out/final/bench_cfg5_sorted.json 2>/dev/null
timeout 300 python bench.py --impl reference --steps 60 --warmup 3 > gpurun_out/final/bench_reference.json 2>/dev/null
ls -la gpurun_out/final
