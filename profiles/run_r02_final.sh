#!/bin/bash
# round 2, final pass on one GPU: tests, smoke, the driver's two bench commands, launch list of the same command, full ncu
O=gpurun_out
(nvidia-smi topo -m; lscpu | head -25; cat /sys/devices/system/node/online; for d in /sys/bus/pci/devices/*; do if grep -qi 0x10de $d/vendor 2>/dev/null; then echo $d $(cat $d/numa_node) $(cat $d/local_cpulist); fi; done) > $O/r02z_topology.txt 2>&1
python -m pytest tests -m gpu -q > $O/r02z_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02z_pytest.log
tail -3 $O/r02z_pytest.log
python __graft_entry__.py smoke > $O/r02z_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r02z_smoke.log; tail -2 $O/r02z_smoke.log
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/r02z_bench_reference.json 2> $O/r02z_bench_reference.err
python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r02z_bench_n1.json 2> $O/r02z_bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"yb::" -c 600 --csv --log-file $O/r02z_launches.csv python bench.py --only --steps 20 --warmup 5 > $O/r02z_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"yb::" -c 40 -o $O/r02z_kernels -f python profiles/ncu_target_r02.py > $O/r02z_ncu_full.log 2>&1
python profiles/summarize_launches.py $O/r02z_launches.csv | tail -8
python - <<PY
import json
d=json.loads(open('$O/r02z_bench_n1.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'warmup', d['warmup'], 'floor', round(d['step_floor_frac'],3), 'roofline', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value']), 'cpu', round(d['cpu_baseline']['value'],1))
r=json.loads(open('$O/r02z_bench_reference.json').read().strip().splitlines()[-1])
print('reference arm', r['value'], r['cpu_baseline']['kind'])
PY
# pipeline depth on the big-batch configs (cfg 3: 887 MB per batch, cfg 5: 5.6 GB)
rm -f $O/r02z_depth.jsonl
for D in 2 3 4; do
  python bench.py --only --workload tiny-416 --batch 1024 --steps 200 --warmup 10 --depth $D >> $O/r02z_depth.jsonl 2>> $O/r02z_depth.err
  python bench.py --only --workload spp-1024 --batch 256 --steps 60 --warmup 5 --depth $D >> $O/r02z_depth.jsonl 2>> $O/r02z_depth.err
done
python -c "
import json
for l in open('$O/r02z_depth.jsonl'):
    d=json.loads(l); print(d['config']['workload'][:10], 'depth', d['batches_in_flight'], round(d['ms_per_step']*1e3,1),'us', round(d['step_floor_frac'],3))
"
