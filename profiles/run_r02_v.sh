#!/bin/bash
# round 2, pass v: fused head reads 19x19 / 13x13 planes in place (loader warps, 4-byte cp.async into the swizzled layout)
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_head.py -m gpu -x -q > $O/r02v_pytest_head.log 2>&1; echo "pytest rc=$?" >> $O/r02v_pytest_head.log
tail -15 $O/r02v_pytest_head.log
python __graft_entry__.py smoke 2>&1 | tail -1
for wl in "--workload spp-608" "--workload tiny-416 --batch 1024"; do
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-configs --no-drop-in $wl > $O/r02v_bench_head.json 2> $O/r02v_bench_head.err
  python -c "
import json
d=json.loads(open('$O/r02v_bench_head.json').read().strip().splitlines()[-1])
h=d['head_fusion']
print(d['config']['workload'][:12], 'fused', round(h['fused_us'],1), 'with pad', h['fused_with_padded_copy_us'], 'unfused', round(h['unfused_us'],1), 'frac', round(h['frac_of_hbm_peak'],3), 'tflops', round(h['tf32_tflops']), 'x3', round(h['fp32x3']['fused_us'],1), 'loader', h['loader_warp_scales'], 'pipe', round(h['pipeline']['ms_per_step'],4))
"
  cp $O/r02v_bench_head.json $O/r02v_bench_head_$(echo $wl | cut -c12-19 | tr -d ' ').json
done
