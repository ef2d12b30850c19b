#!/bin/bash
# round 2, pass y15: shared-memory carve-out preference of the pipelined kernels (an SM changes its L1 / shared split only when idle):
# unset (driver's choice per kernel) vs the same preference for every kernel; and the finalize key-area size (20 vs 24 KB per CTA)
O=gpurun_out
rm -f $O/r02${TAG:-y15}.txt
for c in ${CARVES:-unset 100 50 25 0}; do
  for wl in "--workload tiny-416 --batch 1024" "" "--conf 0.001"; do
    echo "carveout=$c $wl" >> $O/r02${TAG:-y15}.txt
    if [ "$c" = "unset" ]; then unset YOLO_B200_CARVEOUT; else export YOLO_B200_CARVEOUT=$c; fi
    python bench.py --only --steps 300 --warmup 20 $wl 2>> $O/r02${TAG:-y15}.err | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('   step', round(d['ms_per_step']*1e3,1), 'us  decode alone', round(d['roofline']['kernel_ms']*1e3,1), 'floor', round(d['step_floor_frac'],3))" >> $O/r02${TAG:-y15}.txt
  done
done
cat $O/r02${TAG:-y15}.txt
