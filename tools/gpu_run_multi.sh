mkdir -p gpurun_out/multi
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --no-cpu-baseline > gpurun_out/multi/bench_n$N.json 2> gpurun_out/multi/bench_n$N.err; echo "exit $?" >> gpurun_out/multi/bench_n$N.err
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/multi/bench_n1.json 2> gpurun_out/multi/bench_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --impl reference --steps 30 --warmup 3 > gpurun_out/multi/ref_n$N.json 2> gpurun_out/multi/ref_n$N.err
tail -3 gpurun_out/multi/*.err
