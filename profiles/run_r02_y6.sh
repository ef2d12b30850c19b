#!/bin/bash
# round 2, pass y6: in-flight bytes in shared memory (TMA ring variant, 9 warps per SM) instead of registers (LDG variant, 64 K
# registers per SM): the pipelined step on a shape whose planes are all tensor-map describable (spp-640: 20x20, 40x40, 80x80)
O=gpurun_out
rm -f $O/r02y6_steps.jsonl
for v in ldg tma2d; do
  python bench.py --only --workload spp-640 --batch 64 --steps 300 --warmup 20 --variant $v >> $O/r02y6_steps.jsonl 2>> $O/r02y6_steps.err
  python bench.py --only --workload spp-640 --batch 64 --steps 300 --warmup 20 --variant $v --seg-warps 32 >> $O/r02y6_steps.jsonl 2>> $O/r02y6_steps.err
done
python -c "
import json
for l in open('$O/r02y6_steps.jsonl'):
    d=json.loads(l); print(d['config']['workload'][:10], d['roofline']['variant'], 'kernel alone', round(d['roofline']['kernel_ms']*1e3,1), 'step', round(d['ms_per_step']*1e3,1),'us', round(d['step_floor_frac'],3))
"
tail -3 $O/r02y6_steps.err
