#!/bin/bash
# round 2, pass p: fused head with the padded copy overlapped (side stream) with the kernel of the other scales
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_head.py -m gpu -x -q > $O/r02p_pytest_head.log 2>&1; echo "pytest rc=$?" >> $O/r02p_pytest_head.log
tail -4 $O/r02p_pytest_head.log
for wl in "--workload spp-608" "--workload tiny-416 --batch 1024"; do
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-configs --no-drop-in $wl > $O/r02p_bench_head.json 2> $O/r02p_bench_head.err
  python -c "
import json
d=json.loads(open('$O/r02p_bench_head.json').read().strip().splitlines()[-1])
h=d['head_fusion']
print(d['config']['workload'][:12], 'fused', round(h['fused_us'],1), 'serial pad', round(h['fused_pad_not_overlapped_us'],1), 'unfused', round(h['unfused_us'],1), 'frac', round(h['frac_of_hbm_peak'],3), 'tflops', round(h['tf32_tflops']), 'x3', round(h['fp32x3']['fused_us'],1), 'pipe', round(h['pipeline']['ms_per_step'],4))
"
  cp $O/r02p_bench_head.json $O/r02p_bench_head_$(echo $wl | cut -c12-19 | tr -d ' ').json
done
