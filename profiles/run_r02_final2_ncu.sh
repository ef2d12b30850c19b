#!/bin/bash
# round 2, final pass after the packed-group NMS (ncu part): launch list of the driver's command and a --set full capture of the
# candidate stage + NMS kernels on three configurations (summarised on the box: the report itself stays out of gpurun_out when big)
O=gpurun_out
K='regex:decode_|nms_|bucket_|compact_from|pad_planes|build_targets|flag_|scale_'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 600 --csv --log-file $O/r02zz_launches.csv python bench.py --only --steps 20 --warmup 5 > $O/r02zz_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k "$K" -c 20 -o /tmp/r02zz_nms -f python profiles/ncu_target_r02_nms.py > $O/r02zz_ncu_full.log 2>&1
python profiles/ncu_raw_summary.py /tmp/r02zz_nms.ncu-rep > $O/r02zz_ncu_full_summary.txt 2>&1
ls -la /tmp/r02zz_nms.ncu-rep
if [ $(stat -c %s /tmp/r02zz_nms.ncu-rep) -lt 30000000 ]; then cp /tmp/r02zz_nms.ncu-rep $O/; fi
python profiles/summarize_launches.py $O/r02zz_launches.csv | tail -8
grep -c "^==" $O/r02zz_ncu_full_summary.txt
