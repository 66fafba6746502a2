#!/bin/bash
# round 2, 8-GPU session: host-link ceiling with 1/2/4/8 ranks, bench --gpus 8 (configs[3], configs[4] extras inside)
set -x
O=gpurun_out/r2m8
mkdir -p $O
nvidia-smi -L > $O/gpus.txt
nvidia-smi topo -m > $O/topo.txt 2>&1
for N in 1 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N tools/pcie_bench.py > $O/pcie_n$N.json 2> $O/pcie_n$N.err
  tail -1 $O/pcie_n$N.json
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 5 > $O/bench_n8.json 2> $O/bench_n8.err
tail -3 $O/bench_n8.err
( time python -m pytest tests/test_gpu_multi.py -m gpu -q -x ) > $O/pytest_multi8.log 2>&1
tail -3 $O/pytest_multi8.log
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2m8/bench_n8.json').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('value', 'ms_per_step', 'n_gpus')}, d['roofline']['frac'], d['e2e'])
print(json.dumps(d['other_configs'])[:2000])
PY
