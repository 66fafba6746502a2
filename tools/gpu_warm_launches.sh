# per-kernel durations with warm caches (ncu does not flush between launches): usage gpu_warm_launches.sh <bench args>
mkdir -p gpurun_out/warm
CMD="python bench.py $@ --no-graph --steps 20 --warmup 5 --no-e2e --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/warm/plain.json 2>/dev/null && timeout 600 ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none -c 120 --csv --log-file gpurun_out/warm/launches.csv $CMD > /dev/null 2>&1
grep "w2a::" gpurun_out/warm/launches.csv | awk -F'","' '{split($5,a,"("); print a[1], $NF}' | tail -12
