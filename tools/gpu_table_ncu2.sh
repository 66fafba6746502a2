mkdir -p gpurun_out/table
CMD="python bench.py --variant table --no-graph --steps 12 --warmup 4 --no-e2e --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/table/plain.json 2>/dev/null && timeout 900 ncu --set full --cache-control none --clock-control none --import-source on -k regex:logit_table -s 4 -c 1 -o gpurun_out/table/prof_table4 $CMD > gpurun_out/table/ncu_full4.log 2>&1
ls -la gpurun_out/table
