#!/bin/bash
# round 2, 2-GPU session: NCCL determinism test (1 vs 2 GPUs), bench --gpus 2
set -x
O=gpurun_out/r2m2
mkdir -p $O
nvidia-smi -L > $O/gpus.txt
( time python -m pytest tests/test_gpu_multi.py -m gpu -q -x ) > $O/pytest_multi.log 2>&1
tail -5 $O/pytest_multi.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err
tail -3 $O/bench_n2.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2m2/bench_n2.json').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('value', 'ms_per_step', 'n_gpus')}, d['roofline']['frac'], d['e2e']['value'], d['e2e'].get('link_gbs'), d['e2e'].get('frac_of_link'))
print(json.dumps(d['other_configs'])[:1500])
PY
