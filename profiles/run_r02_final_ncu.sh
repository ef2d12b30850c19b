#!/bin/bash
# round 2, final pass (ncu part): launch list of the driver's command and a --set full capture of every kernel
O=gpurun_out
K='regex:decode_|nms_|bucket_|compact_from|pad_planes|build_targets|flag_|scale_'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 600 --csv --log-file $O/r02z_launches.csv python bench.py --only --steps 20 --warmup 5 > $O/r02z_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k "$K" -c 40 -o $O/r02z_kernels -f python profiles/ncu_target_r02.py > $O/r02z_ncu_full.log 2>&1
python profiles/summarize_launches.py $O/r02z_launches.csv | tail -8
tail -2 $O/r02z_ncu_full.log
