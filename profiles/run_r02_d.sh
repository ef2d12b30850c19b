#!/bin/bash
# round 2, pass d: ncu of the NMS kernels at conf 0.001 and 0.3 (launch list + full capture of segment / finalize)
O=gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"nms_|bucket_|decode_compact" -c 40 --csv --log-file $O/r02d_launches_0.001.csv python profiles/bench_kernels.py spp-608 64 0.001 > $O/r02d_ncu1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"nms_|bucket_|decode_compact" -c 40 --csv --log-file $O/r02d_launches_0.3.csv python profiles/bench_kernels.py spp-608 64 0.3 > $O/r02d_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"nms_segment|nms_finalize" -s 4 -c 2 -o $O/r02d_nms_0.001 -f python profiles/bench_kernels.py spp-608 64 0.001 > $O/r02d_ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"nms_segment|nms_finalize|bucket" -s 6 -c 3 -o $O/r02d_nms_0.3 -f python profiles/bench_kernels.py spp-608 64 0.3 > $O/r02d_ncu4.log 2>&1
python profiles/summarize_launches.py $O/r02d_launches_0.001.csv | tail -6
python profiles/summarize_launches.py $O/r02d_launches_0.3.csv | tail -6
