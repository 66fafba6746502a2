mkdir -p gpurun_out/ko; rm -f gpurun_out/ko/*
B="python bench.py --no-e2e --no-cpu-baseline --variant coop --steps 1500 --warmup 20"
for ko in 0 127 128 255 256 383 511 384; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -shared -DW2A_KO=$ko -I include -I tools/exp -o /tmp/ko_$ko.so tools/exp/w2a.cu 2>/dev/null
  W2A_LIB_PATH=/tmp/ko_$ko.so timeout 200 $B 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('KO=$ko ms/step %.4f'%d['ms_per_step'])" >> gpurun_out/ko/results.txt
done
cat gpurun_out/ko/results.txt
