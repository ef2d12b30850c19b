#!/bin/bash
# round 2, pass y7: loads in flight per thread on the scalar path (planes that are not a multiple of 4 floats: 19x19, 13x13)

O=gpurun_out
python profiles/tune_step.py --only --steps 300 --warmup 20 > $O/r02y7_cfg2.txt 2>&1
python profiles/tune_step.py --only --steps 300 --warmup 20 --workload tiny-416 --batch 1024 > $O/r02y7_cfg3.txt 2>&1
cat $O/r02y7_cfg2.txt $O/r02y7_cfg3.txt
