#!/bin/bash
# round 2, pass c: warp-CTA segment kernel + register finalize sort: parity, then cfg 2 / cfg 4 step times and launch lists
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r02c_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02c_pytest.log
tail -5 $O/r02c_pytest.log
for extra in "--depth 3" "--depth 6" "--depth 3 --conf 0.001" "--depth 6 --conf 0.001"; do
  python bench.py --only --steps 1000 --warmup 20 $extra >> $O/r02c_steps.jsonl 2>> $O/r02c_steps.err
done
python profiles/bench_kernels.py spp-608 64 0.3 > $O/r02c_kernels_0.3.json 2>&1
python profiles/bench_kernels.py spp-608 64 0.001 > $O/r02c_kernels_0.001.json 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/r02c_launches_0.001.csv python bench.py --only --steps 3 --warmup 3 --conf 0.001 --no-graph > $O/r02c_ncu1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/r02c_launches_0.3.csv python bench.py --only --steps 3 --warmup 3 --no-graph > $O/r02c_ncu2.log 2>&1
python -c "
import json
for l in open('$O/r02c_steps.jsonl'):
    d=json.loads(l); print(d['config']['conf_thres'], d['batches_in_flight'], round(d['ms_per_step']*1e3,1),'us', d['step_floor_frac'])
"
