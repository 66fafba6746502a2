mkdir -p gpurun_out/v
rm -f gpurun_out/v/*.json
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
B="python bench.py --no-e2e --no-cpu-baseline --variant coop"
for mb in 2 3 4; do for w in 4 8; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -DW2A_COOP_MINBLOCKS=$mb -DW2A_COOP_WARPS=$w -I include -o /tmp/lib_${mb}_${w}.so weather2alert_b200/csrc/w2a.cu 2>/dev/null
  W2A_LIB_PATH=/tmp/lib_${mb}_${w}.so timeout 200 $B > gpurun_out/v/cfg3_mb${mb}_w${w}.json 2>/dev/null
done; done
timeout 300 python bench.py > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err
timeout 300 $B --workload cfg4_2m > gpurun_out/bench_cfg4.json 2>/dev/null
timeout 300 $B --workload cfg2_4096 > gpurun_out/bench_cfg2_coop.json 2>/dev/null
CMD="python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_coop -s 8 -c 2 -o gpurun_out/prof_coop_v3 $CMD > gpurun_out/ncu_full.log 2>&1
ls gpurun_out gpurun_out/v
