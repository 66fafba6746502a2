# usage: gpu_ko_table.sh ko...   knock-out study of step_table_kernel (results are wrong by design; timing only)
B="python bench.py --no-e2e --no-cpu-baseline --variant table"
for c in $@; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -DW2A_TABLE_KO=$c -I include -o /tmp/lib_ko$c.so weather2alert_b200/csrc/w2a.cu 2>/dev/null
  W2A_LIB_PATH=/tmp/lib_ko$c.so timeout 200 $B 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('KO=$c', 'ms/step %.4f'%d['ms_per_step'])"
done
