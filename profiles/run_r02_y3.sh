#!/bin/bash
# round 2, pass y3: decode CTA granularity (32 / 64 / 128 / 256 threads at the same 512 threads per SM) vs the pipelined step:
# a single-warp NMS CTA (2 K registers) keeps a whole decode CTA's registers (16 K at 128 threads) from being re-used
O=gpurun_out
python profiles/tune_step.py --only --steps 300 --warmup 20 > $O/r02y3_cfg2.txt 2>&1
python profiles/tune_step.py --only --steps 300 --warmup 20 --workload tiny-416 --batch 1024 > $O/r02y3_cfg3.txt 2>&1
python profiles/tune_step.py --only --steps 300 --warmup 20 --conf 0.001 > $O/r02y3_cfg4.txt 2>&1
tail -5 $O/r02y3_cfg2.txt $O/r02y3_cfg3.txt $O/r02y3_cfg4.txt
