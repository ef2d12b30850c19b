#!/bin/bash
# round 2, pass n: segment kernel prefetches the offsets of a chunk of pairs; parity + cfg 3 / 2 / 4 / 5 steps
O=gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_random.py tests/test_gpu_fullsize.py tests/test_gpu_pipeline.py -m gpu -x -q > $O/r02n_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02n_pytest.log
tail -3 $O/r02n_pytest.log
rm -f $O/r02n_steps.jsonl
python bench.py --only --workload tiny-416 --batch 1024 --steps 300 --warmup 20 >> $O/r02n_steps.jsonl 2>> $O/r02n_steps.err
python bench.py --only --workload tiny-416 --batch 1024 --steps 30 --warmup 5 >> $O/r02n_steps.jsonl 2>> $O/r02n_steps.err
python bench.py --only --steps 300 --warmup 20 >> $O/r02n_steps.jsonl 2>> $O/r02n_steps.err
python bench.py --only --steps 300 --warmup 20 --conf 0.001 >> $O/r02n_steps.jsonl 2>> $O/r02n_steps.err
python bench.py --only --workload spp-1024 --batch 256 --steps 100 --warmup 10 >> $O/r02n_steps.jsonl 2>> $O/r02n_steps.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"nms_|bucket_|decode_compact" -c 16 --csv --log-file $O/r02n_launches_tiny.csv python profiles/bench_kernels.py tiny-416 1024 0.3 > $O/r02n_ncu1.log 2>&1
python profiles/summarize_launches.py $O/r02n_launches_tiny.csv | tail -5
python -c "
import json
for l in open('$O/r02n_steps.jsonl'):
    d=json.loads(l); print(d['config']['workload'][:10], d['config']['conf_thres'], d['steps'], round(d['ms_per_step']*1e3,1),'us', round(d['step_floor_frac'],3))
"
