#!/bin/bash
# round 2, pass y11: decode_compact with its loads in flight parked in shared memory (cp.async) instead of registers:
# parity with the variant library, then the pipelined step for 4 .. 7 resident CTAs per SM
O=gpurun_out
YOLO_B200_LIB=$PWD/build/variants/st4.so python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_random.py -m gpu -x -q > $O/r02y11_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02y11_pytest.log
tail -3 $O/r02y11_pytest.log
python profiles/tune_step.py --only --steps 300 --warmup 20 > $O/r02y11_cfg2.txt 2>&1
python profiles/tune_step.py --only --steps 300 --warmup 20 --workload tiny-416 --batch 1024 > $O/r02y11_cfg3.txt 2>&1
cat $O/r02y11_cfg2.txt $O/r02y11_cfg3.txt
