#!/bin/bash
# round 2, pass y: small segments packed several to a warp (nms_packed_group): parity suite, pipelined steps, per-launch times
O=gpurun_out
T=${1:-y}
python -m pytest tests/test_gpu_parity.py tests/test_gpu_random.py tests/test_gpu_fullsize.py tests/test_gpu_pipeline.py -m gpu -x -q > $O/r02${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02${T}_pytest.log
tail -3 $O/r02${T}_pytest.log
rm -f $O/r02${T}_steps.jsonl
python bench.py --only --steps 300 --warmup 20 >> $O/r02${T}_steps.jsonl 2>> $O/r02${T}_steps.err
python bench.py --only --steps 20 --warmup 5 >> $O/r02${T}_steps.jsonl 2>> $O/r02${T}_steps.err
python bench.py --only --steps 300 --warmup 20 --conf 0.001 >> $O/r02${T}_steps.jsonl 2>> $O/r02${T}_steps.err
python bench.py --only --workload tiny-416 --batch 1024 --steps 300 --warmup 20 >> $O/r02${T}_steps.jsonl 2>> $O/r02${T}_steps.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"nms_|bucket_" -c 12 --csv --log-file $O/r02${T}_launches_0.3.csv python profiles/bench_kernels.py spp-608 64 0.3 > $O/r02${T}_ncu2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"nms_|bucket_" -c 12 --csv --log-file $O/r02${T}_launches_tiny.csv python profiles/bench_kernels.py tiny-416 1024 0.3 > $O/r02${T}_ncu3.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"nms_|bucket_" -c 12 --csv --log-file $O/r02${T}_launches_0.001.csv python profiles/bench_kernels.py spp-608 64 0.001 > $O/r02${T}_ncu1.log 2>&1
python profiles/summarize_launches.py $O/r02${T}_launches_0.3.csv | tail -4
python profiles/summarize_launches.py $O/r02${T}_launches_tiny.csv | tail -4
python profiles/summarize_launches.py $O/r02${T}_launches_0.001.csv | tail -4
python -c "
import json
for l in open('$O/r02${T}_steps.jsonl'):
    d=json.loads(l); print(d['config']['workload'][:10], d['config']['conf_thres'], d['steps'], round(d['ms_per_step']*1e3,1),'us', round(d['step_floor_frac'],3))
"
