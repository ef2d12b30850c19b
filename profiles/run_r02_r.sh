#!/bin/bash
# round 2, pass r: 2-GPU steady state vs pipeline depth (where do the ~7-12 us/step over N = 1 come from?)
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544"
rm -f $O/r02r_steps.jsonl
for extra in "--steps 300 --depth 3" "--steps 300 --depth 4" "--steps 300 --depth 6" "--steps 20 --depth 4" "--steps 20 --depth 6"; do
  timeout 300 $TR bench.py --gpus 2 --only --warmup 20 $extra 2>> $O/r02r_steps.err | grep "^{" >> $O/r02r_steps.jsonl
done
python bench.py --only --steps 300 --warmup 20 >> $O/r02r_steps.jsonl 2>> $O/r02r_steps.err
python -c "
import json
for l in open('$O/r02r_steps.jsonl'):
    d=json.loads(l); print('N', d['n_gpus'], 'steps', d['steps'], 'depth', d['batches_in_flight'], [round(x*1e3,1) for x in d['ms_per_step_by_rank']], d.get('gather_check'))
"
