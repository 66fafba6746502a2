# usage: gpu_sweep.sh <tag> [configs...]   each config = "minblocks:warps:ahead"   (run on the GPU box through gpurun)
TAG=${1:-x}; shift
CFGS=${@:-"3:8:2"}
mkdir -p gpurun_out/v
rm -f gpurun_out/v/*.json
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
B="python bench.py --no-e2e --no-cpu-baseline --variant coop"
for c in $CFGS; do
  IFS=: read mb w ah <<< "$c"
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -DW2A_COOP_MINBLOCKS=$mb -DW2A_COOP_WARPS=$w -DW2A_COOP_AHEAD=$ah -I include -o /tmp/lib_$c.so weather2alert_b200/csrc/w2a.cu 2>/dev/null
  W2A_LIB_PATH=/tmp/lib_$c.so timeout 200 $B > gpurun_out/v/cfg3_mb${mb}_w${w}_a${ah}.json 2>/dev/null
done
timeout 300 python bench.py > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err
timeout 300 $B --workload cfg4_2m > gpurun_out/bench_cfg4.json 2>/dev/null
timeout 300 $B --workload cfg2_4096 > gpurun_out/bench_cfg2_coop.json 2>/dev/null
CMD="python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_coop -s 8 -c 1 -o gpurun_out/prof_coop_$TAG $CMD > gpurun_out/ncu_full.log 2>&1
