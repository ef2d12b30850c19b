#!/bin/bash
# round 2, pass y10: fused head, X / X+W stream rate WITHOUT tensor-core work (profiling flag 0x800)
O=gpurun_out
python profiles/head_probe.py xrate > $O/r02y10_xrate.txt 2>&1
cat $O/r02y10_xrate.txt
