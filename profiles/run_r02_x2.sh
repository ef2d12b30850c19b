#!/bin/bash
# round 2, pass x2: cfg 3 (tiny-416 batch 1024): fewer resident segment warps, no stream priority, TMA decode variant
O=gpurun_out
rm -f $O/r02x2_steps.jsonl
for extra in "--seg-warps 8" "--seg-warps 4" "--no-priority" "--variant tma2d" "--variant tma2d --seg-warps 32"; do
  python bench.py --only --workload tiny-416 --batch 1024 --steps 300 --warmup 20 $extra >> $O/r02x2_steps.jsonl 2>> $O/r02x2_steps.err
done
python -c "
import json
for l in open('$O/r02x2_steps.jsonl'):
    d=json.loads(l); print(d['config']['workload'][:10], d['config']['conf_thres'], d['steps'], 'depth', d['batches_in_flight'], d['roofline']['variant'], round(d['roofline']['kernel_ms']*1e3,1), round(d['ms_per_step']*1e3,1),'us', round(d['step_floor_frac'],3))
"
