#!/bin/bash
# round 2, pass k: does the 20-step figure depend on how long the GPU has been busy before the timed region?
O=gpurun_out
rm -f $O/r02k_warm.jsonl
for W in 5 50 500 3000; do
  for rep in 1 2; do
    python bench.py --only --steps 20 --warmup $W >> $O/r02k_warm.jsonl 2>> $O/r02k_warm.err
  done
done
python bench.py --only --steps 200 --warmup 5 >> $O/r02k_warm.jsonl 2>> $O/r02k_warm.err
python -c "
import json
for l in open('$O/r02k_warm.jsonl'):
    d=json.loads(l); print('steps', d['steps'], 'warmup', d['warmup'], round(d['ms_per_step']*1e3,1),'us', d['clocks'])
"
timeout 600 python -m pytest tests/test_gpu_head.py -m gpu -x -q 2>&1 | tail -3
