#!/bin/bash
# round 2, pass w: N-GPU steady state vs 20-step runs (where does the N = 8 efficiency go?)
N=${1:-8}
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555"
rm -f $O/r02w_steps_n$N.jsonl
nproc > $O/r02w_host_n$N.txt; free -g >> $O/r02w_host_n$N.txt; nvidia-smi topo -m >> $O/r02w_host_n$N.txt 2>&1
for extra in "--steps 300" "--steps 20" "--steps 20" "--steps 20" "--steps 20 --depth 4" "--steps 20 --depth 6"; do
  timeout 200 $TR bench.py --gpus $N --only --warmup 5 $extra 2>> $O/r02w_steps_n$N.err | grep "^{" >> $O/r02w_steps_n$N.jsonl
done
python -c "
import json
for l in open('$O/r02w_steps_n$N.jsonl'):
    d=json.loads(l); print('N', d['n_gpus'], 'steps', d['steps'], 'depth', d['batches_in_flight'], [round(x*1e3,1) for x in d['ms_per_step_by_rank']], d.get('gather_check'))
"
head -1 $O/r02w_host_n$N.txt
