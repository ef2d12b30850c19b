#!/bin/bash
# round 2, pass y13: 128-thread finalize CTAs for small images as the default (generic network in shared memory beyond 2048 rows)
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r02y13_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02y13_pytest.log
tail -2 $O/r02y13_pytest.log
rm -f $O/r02y13_steps.jsonl
python bench.py --only --workload tiny-416 --batch 1024 --steps 300 --warmup 20 >> $O/r02y13_steps.jsonl 2>> $O/r02y13_steps.err
python bench.py --only --steps 300 --warmup 20 >> $O/r02y13_steps.jsonl 2>> $O/r02y13_steps.err
python -c "
import json
for l in open('$O/r02y13_steps.jsonl'):
    d=json.loads(l); print(d['config']['workload'][:10], d['config']['conf_thres'], d['steps'], round(d['ms_per_step']*1e3,1),'us', round(d['step_floor_frac'],3))
"
