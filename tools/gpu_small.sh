for v in thread coop; do
python bench.py --workload cfg2_4096 --no-cpu-baseline --no-e2e --variant $v > gpurun_out/cfg2_graph_$v.json 2> gpurun_out/cfg2_graph_$v.err
python bench.py --workload cfg2_4096 --no-cpu-baseline --no-e2e --variant $v --no-graph > gpurun_out/cfg2_nograph_$v.json 2>/dev/null
done
python bench.py --no-cpu-baseline --no-e2e --graph > gpurun_out/cfg3_graph.json 2> gpurun_out/cfg3_graph.err
python bench.py --no-cpu-baseline --no-e2e --envs-per-gpu 65536 > gpurun_out/n64k.json 2>/dev/null
python bench.py --no-cpu-baseline --no-e2e --envs-per-gpu 262144 > gpurun_out/n256k.json 2>/dev/null
for f in cfg2_graph_thread cfg2_nograph_thread cfg2_graph_coop cfg2_nograph_coop cfg3_graph n64k n256k; do python -c "
import json
d=json.load(open('gpurun_out/$f.json')); print('$f', '%.3e'%d['value'], 'ms/step %.5f'%d['ms_per_step'], d['steps'], d['gpu_launches'], d['config']['mode'], 'rollout %.3e'%d['rollout']['value'], d['rollout']['same_returns_as_stepping'])
"; done; tail -3 gpurun_out/cfg2_graph_thread.err
