#!/bin/bash
# round 2, pass y18: state after the finalize / carve-out changes: GPU suite, the three configurations, the driver's command, launch list
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r02y18_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02y18_pytest.log
tail -2 $O/r02y18_pytest.log
CARVES="unset" TAG=y18 bash profiles/run_r02_y15.sh
python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r02y18_bench_n1.json 2> $O/r02y18_bench_n1.err
K='regex:decode_|nms_|bucket_|compact_from|pad_planes|build_targets|flag_|scale_'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 600 --csv --log-file $O/r02y18_launches.csv python bench.py --only --steps 20 --warmup 5 > $O/r02y18_ncu_launches.log 2>&1
python profiles/summarize_launches.py $O/r02y18_launches.csv | tail -6
python - <<PY
import json
d=json.loads(open('$O/r02y18_bench_n1.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'floor', round(d['step_floor_frac'],3), 'roofline', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value']), 'cpu', round(d['cpu_baseline']['value'],1))
for c in d['configs']: print(c['config'][:30], round(c.get('ms_per_step',0)*1e3,1), c.get('step_floor_frac'))
PY
