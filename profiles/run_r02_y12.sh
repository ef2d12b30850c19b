#!/bin/bash
# round 2, pass y12: finalize CTA size for images that cannot stage many rows (256 -> 128 / 64 threads): cfg 3 step, parity
O=gpurun_out
for v in fin128 fin64; do
  YOLO_B200_LIB=$PWD/build/variants/$v.so python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_random.py -m gpu -x -q > $O/r02y12_pytest_$v.log 2>&1; echo "pytest $v rc=$?"; tail -1 $O/r02y12_pytest_$v.log
done
python profiles/tune_step.py --only --steps 300 --warmup 20 --workload tiny-416 --batch 1024 > $O/r02y12_cfg3.txt 2>&1
python profiles/tune_step.py --only --steps 300 --warmup 20 --workload tiny-416 --batch 1024 >> $O/r02y12_cfg3.txt 2>&1
cat $O/r02y12_cfg3.txt
