# usage: gpu_sweep_table.sh minblocks...   (step_table_kernel launch bound)
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
B="python bench.py --no-e2e --no-cpu-baseline --variant table"
for c in $@; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -DW2A_TABLE_MINBLOCKS=$c -I include -o /tmp/lib_$c.so weather2alert_b200/csrc/w2a.cu 2>/dev/null
  for loc in interleaved sorted; do
  W2A_LIB_PATH=/tmp/lib_$c.so timeout 200 $B --locality $loc 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$c $loc', '%.3e'%d['value'], 'ms/step %.4f'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'])"
  done
done
