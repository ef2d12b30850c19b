#!/bin/bash
# round 2, pass g: ballot-ranked radix finalize, 1024-thread bucket kernel: parity + timings
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r02g_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02g_pytest.log
tail -4 $O/r02g_pytest.log
rm -f $O/r02g_steps.jsonl
for extra in "--depth 3" "--depth 4" "--depth 6" "--depth 3 --conf 0.001" "--depth 6 --conf 0.001"; do
  python bench.py --only --steps 1000 --warmup 20 $extra >> $O/r02g_steps.jsonl 2>> $O/r02g_steps.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"nms_|bucket_" -c 12 --csv --log-file $O/r02g_launches_0.001.csv python profiles/bench_kernels.py spp-608 64 0.001 > $O/r02g_ncu1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"nms_|bucket_" -c 12 --csv --log-file $O/r02g_launches_0.3.csv python profiles/bench_kernels.py spp-608 64 0.3 > $O/r02g_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"nms_segment|nms_finalize|bucket" -s 6 -c 3 -o $O/r02g_nms_0.001 -f python profiles/bench_kernels.py spp-608 64 0.001 > $O/r02g_ncu3.log 2>&1
python profiles/summarize_launches.py $O/r02g_launches_0.001.csv | tail -4
python profiles/summarize_launches.py $O/r02g_launches_0.3.csv | tail -4
python -c "
import json
for l in open('$O/r02g_steps.jsonl'):
    d=json.loads(l); print(d['config']['conf_thres'], d['batches_in_flight'], round(d['ms_per_step']*1e3,1),'us', round(d['step_floor_frac'],3))
"
