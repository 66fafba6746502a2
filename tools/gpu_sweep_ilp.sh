# usage: gpu_sweep_ilp.sh ilp...   warm per-kernel times of the TABLE variant for W2A_TABLE_ILP values
for c in $@; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -DW2A_TABLE_ILP=$c -I include -o /tmp/lib_ilp$c.so weather2alert_b200/csrc/w2a.cu 2>/dev/null
  echo "ILP=$c"; W2A_LIB_PATH=/tmp/lib_ilp$c.so bash tools/gpu_warm_launches.sh --variant table | tail -2
  W2A_LIB_PATH=/tmp/lib_ilp$c.so python bench.py --variant table --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ms/step %.4f'%d['ms_per_step'])"
done
