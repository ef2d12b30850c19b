#!/bin/bash
# round 2, pass o: segment-kernel residency on cfg 3 (tiny-416 batch 1024: 82 k (image, class) pairs, ~4 boxes each)
O=gpurun_out
rm -f $O/r02o_steps.jsonl
for R in 16 24 32; do
  echo "R=$R" >> $O/r02o_steps.jsonl
  python bench.py --only --workload tiny-416 --batch 1024 --steps 300 --warmup 20 --seg-warps $R >> $O/r02o_steps.jsonl 2>> $O/r02o_steps.err
done
python -c "
import json
for l in open('$O/r02o_steps.jsonl'):
    if l.startswith('R='): print(l.strip(), end=' -> '); continue
    d=json.loads(l); print(d['config']['workload'][:10], round(d['ms_per_step']*1e3,1),'us', round(d['step_floor_frac'],3))
"
