# usage: gpu_sweep2.sh cfg...   each cfg = minblocks:warps:ahead:pre
mkdir -p gpurun_out/v2; rm -f gpurun_out/v2/*
B="python bench.py --no-e2e --no-cpu-baseline --variant coop"
for c in $@; do
  IFS=: read mb w ah pre <<< "$c"
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -DW2A_COOP_MINBLOCKS=$mb -DW2A_COOP_WARPS=$w -DW2A_COOP_AHEAD=$ah -DW2A_COOP_PRE=${pre:-2} -I include -o /tmp/lib_$c.so weather2alert_b200/csrc/w2a.cu 2>/dev/null
  for loc in interleaved sorted; do
  W2A_LIB_PATH=/tmp/lib_$c.so timeout 200 $B --locality $loc 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$c $loc', '%.3e'%d['value'], 'ms/step %.4f'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'])"
  done
done
