#!/bin/bash
# round 2, pass y8: A/B on one box, alternating: library with the previous NMS kernels (build/variants/a_head_nms.so) vs the packed
# small-segment version, the driver's command (20 steps) and the steady state (300 steps)
O=gpurun_out
rm -f $O/r02y8.txt
for rep in 1 2 3; do
  python profiles/tune_step.py --only --steps 20 --warmup 5 >> $O/r02y8.txt 2>&1
done
python profiles/tune_step.py --only --steps 300 --warmup 20 >> $O/r02y8.txt 2>&1
python profiles/tune_step.py --only --steps 300 --warmup 20 >> $O/r02y8.txt 2>&1
cat $O/r02y8.txt
