mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err; echo "exit $?" >> gpurun_out/bench_cfg3.err
timeout 300 python bench.py --variant thread --no-e2e --no-cpu-baseline > gpurun_out/bench_cfg3_thread.json 2> gpurun_out/bench_cfg3_thread.err
timeout 300 python bench.py --workload cfg2_4096 --no-cpu-baseline > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err
timeout 300 python bench.py --workload cfg2_4096 --variant coop --no-e2e --no-cpu-baseline > gpurun_out/bench_cfg2_coop.json 2> gpurun_out/bench_cfg2_coop.err
timeout 300 python bench.py --workload cfg4_2m --no-cpu-baseline > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err
timeout 300 python bench.py --workload cfg5_8m --no-cpu-baseline --no-e2e --steps 600 > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err
timeout 400 python bench.py --impl reference --steps 60 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
CMD="python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
timeout 300 $CMD > gpurun_out/plain2.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_coop -s 8 -c 2 -o gpurun_out/prof_coop $CMD > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
