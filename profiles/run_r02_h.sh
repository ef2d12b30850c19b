#!/bin/bash
# round 2, pass h: segment-kernel residency (persistent ticketed grid) vs step time, cfg 2 and cfg 4
O=gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_random.py -m gpu -x -q > $O/r02h_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02h_pytest.log
tail -3 $O/r02h_pytest.log
rm -f $O/r02h_steps.jsonl
for R in 32 24 16 12 8; do
  for extra in "--depth 3" "--depth 3 --conf 0.001"; do
    echo "R=$R $extra" >> $O/r02h_steps.jsonl
    YOLO_B200_NMS_SEG_RESIDENCY=$R python bench.py --only --steps 1000 --warmup 20 $extra >> $O/r02h_steps.jsonl 2>> $O/r02h_steps.err
  done
done
python -c "
import json
for l in open('$O/r02h_steps.jsonl'):
    if l.startswith('R='): print(l.strip(), end=' -> '); continue
    d=json.loads(l); print(d['config']['conf_thres'], d['batches_in_flight'], round(d['ms_per_step']*1e3,1),'us', round(d['step_floor_frac'],3))
"
