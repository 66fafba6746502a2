cd gpurun_out; tail -4 pytest.log; for f in v/*.json bench_cfg3.json bench_cfg4.json bench_cfg2_coop.json; do python -c "
import json
try:
    d=json.load(open('$f')); print('$f', '%.3e'%d['value'], 'ms/step %.4f'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'], 'canon %.3f'%d['roofline_canonical_466B']['frac'], d['clocks']['sm_mhz'], d['clocks']['reasons'], (d.get('e2e') or {}).get('value'))
except Exception as e: print('$f ERR', e)
"; done
