#!/bin/bash
# round 2, pass j: fused head fp32x3 mode (3 TF32 passes): parity against the oracle + timing; full test-suite
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_head.py -m gpu -x -q > $O/r02j_pytest_head.log 2>&1; echo "pytest rc=$?" >> $O/r02j_pytest_head.log
tail -15 $O/r02j_pytest_head.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-configs --no-drop-in > $O/r02j_bench_head.json 2> $O/r02j_bench_head.err
python -c "
import json
d=json.loads(open('$O/r02j_bench_head.json').read().strip().splitlines()[-1])
print(json.dumps(d.get('head_fusion'), indent=1))
print('step', d['ms_per_step'])
"
