mkdir -p gpurun_out/ko
for ko in 511; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -DW2A_KO=$ko -I include -I tools/exp -o /tmp/ko_$ko.so tools/exp/w2a.cu 2>/dev/null
  CMD="python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --variant coop"
  W2A_LIB_PATH=/tmp/ko_$ko.so timeout 300 $CMD > gpurun_out/ko/plain_$ko.log 2>&1 && W2A_LIB_PATH=/tmp/ko_$ko.so timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_coop -s 8 -c 1 -o gpurun_out/ko/prof_ko$ko $CMD > gpurun_out/ko/ncu_$ko.log 2>&1
done
ls -la gpurun_out/ko
