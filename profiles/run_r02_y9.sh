#!/bin/bash
# round 2, pass y9: GPU suite after the NMS change (packed small segments) + fused head: X / X+W stream rate WITHOUT tensor-core
# work (profiling flag 0x800) -- is the 4.6 TB/s ceiling of the X stream a property of the TMA pattern or of the MMA-coupled ring?
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r02y9_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02y9_pytest.log
tail -3 $O/r02y9_pytest.log
python profiles/head_probe.py xrate > $O/r02y9_xrate.txt 2>&1
cat $O/r02y9_xrate.txt
