#!/bin/bash
O=gpurun_out
python profiles/tune_step.py --only --steps 300 --warmup 20 --workload tiny-416 --batch 1024 > $O/r02y14_cfg3.txt 2>&1
python profiles/tune_step.py --only --steps 300 --warmup 20 --workload tiny-416 --batch 1024 >> $O/r02y14_cfg3.txt 2>&1
cat $O/r02y14_cfg3.txt
