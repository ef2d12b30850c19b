#!/bin/bash
# round 2, pass l: full GPU suite (targets kernel, fp32x3, eval loop, dense TMA), smoke, the driver's bench command
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r02l_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02l_pytest.log
tail -5 $O/r02l_pytest.log
python __graft_entry__.py smoke > $O/r02l_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r02l_smoke.log; tail -2 $O/r02l_smoke.log
python bench.py --steps 20 --warmup 5 > $O/r02l_bench_driver.json 2> $O/r02l_bench_driver.err
python - <<PY
import json
d=json.loads(open('$O/r02l_bench_driver.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'warmup', d['warmup'], 'floor', round(d['step_floor_frac'],3), 'roofline', round(d['roofline']['frac'],3))
print('e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'])
print('drop_in', {k: (round(v,3) if isinstance(v,float) else v) for k,v in d['drop_in'].items() if k!='what'})
for c in d['configs']: print(c['config'], round(c.get('value',0)), c.get('ms_per_step'), c.get('step_floor_frac'))
print('head pipeline', d['head_fusion']['pipeline'])
PY
