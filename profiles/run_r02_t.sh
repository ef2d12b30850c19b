#!/bin/bash
# round 2, pass t: the driver's 20-step command at N = 2 and N = 1, repeated (host-side trims: cached root views, end event
# enqueued right after the last submit, start events aligned by an enqueued all-reduce), plus the steady state
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544"
rm -f $O/r02t_steps.jsonl
for extra in "--steps 20" "--steps 20" "--steps 20" "--steps 20" "--steps 300"; do
  timeout 300 $TR bench.py --gpus 2 --only --warmup 5 $extra 2>> $O/r02t_steps.err | grep "^{" >> $O/r02t_steps.jsonl
done
for extra in "--steps 20" "--steps 20" "--steps 20" "--steps 300"; do
  python bench.py --only --warmup 5 $extra >> $O/r02t_steps.jsonl 2>> $O/r02t_steps.err
done
python -m pytest tests/test_gpu_pipeline.py -m gpu -x -q 2>&1 | tail -2
python -c "
import json
for l in open('$O/r02t_steps.jsonl'):
    d=json.loads(l); print('N', d['n_gpus'], 'steps', d['steps'], 'depth', d['batches_in_flight'], [round(x*1e3,1) for x in d['ms_per_step_by_rank']], d.get('gather_check'))
"
