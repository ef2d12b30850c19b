#!/bin/bash
# round 2: the driver's exact bench command at N GPUs (torchrun), plus the 2-rank gather parity check
N=${1:-2}
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
if [ "$N" = "1" ]; then
  python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r02s_bench_n1.json 2> $O/r02s_bench_n1.err
else
  timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 5 > $O/r02s_bench_n$N.json 2> $O/r02s_bench_n$N.err
  timeout 300 $TR tests/multi_gpu_check.py > $O/r02s_multi_gpu_check_n$N.log 2>&1; echo "rc=$?" >> $O/r02s_multi_gpu_check_n$N.log
  tail -2 $O/r02s_multi_gpu_check_n$N.log
fi
python - <<PY
import json
for l in open('$O/r02s_bench_n$N.json'):
    if l.startswith('{'):
        d=json.loads(l); print('N', d['n_gpus'], 'value', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'gather', d.get('gather_check'), [ (c['config'][:5], round(c.get('value',0))) for c in d.get('configs',[]) if 'value' in c])
PY
