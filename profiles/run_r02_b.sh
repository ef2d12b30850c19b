#!/bin/bash
# round 2, 2-GPU pass: gather protocol parity + the driver's bench command at N=2
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tests/multi_gpu_check.py > $O/r02b_multi_gpu_check.log 2>&1; echo "rc=$?" >> $O/r02b_multi_gpu_check.log
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 > $O/r02b_bench_n2_driver.json 2> $O/r02b_bench_n2_driver.err; echo "rc=$?" >> $O/r02b_bench_n2_driver.err
timeout 300 $TR bench.py --gpus 2 --only --steps 2000 --warmup 20 > $O/r02b_bench_n2_long.json 2> $O/r02b_bench_n2_long.err
timeout 300 $TR bench.py --gpus 2 --only --steps 2000 --warmup 20 --depth 3 > $O/r02b_bench_n2_long_d3.json 2> $O/r02b_bench_n2_long_d3.err
python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/r02b_pytest_parity.log 2>&1
tail -3 $O/r02b_multi_gpu_check.log; tail -2 $O/r02b_pytest_parity.log
