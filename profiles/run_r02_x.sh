#!/bin/bash
# round 2, pass x: cfg 3 (tiny-416 batch 1024): residency of the segment kernel and batches in flight vs step time
O=gpurun_out
rm -f $O/r02x_steps.jsonl
for extra in "" "--seg-warps 24" "--seg-warps 32" "--seg-warps 32 --depth 4" "--depth 2"; do
  python bench.py --only --workload tiny-416 --batch 1024 --steps 300 --warmup 20 $extra >> $O/r02x_steps.jsonl 2>> $O/r02x_steps.err
done
python -c "
import json
for l in open('$O/r02x_steps.jsonl'):
    d=json.loads(l); print(d['config']['workload'][:10], d['config']['conf_thres'], d['steps'], 'depth', d['batches_in_flight'], round(d['ms_per_step']*1e3,1),'us', round(d['step_floor_frac'],3))
"
