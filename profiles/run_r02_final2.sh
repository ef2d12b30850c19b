#!/bin/bash
# round 2, final pass (after the packed-group NMS) on one GPU: tests, smoke, the driver's two bench commands, launch list, full ncu
O=gpurun_out
(nvidia-smi topo -m; lscpu | head -25; cat /sys/devices/system/node/online; for d in /sys/bus/pci/devices/*; do if grep -qi 0x10de $d/vendor 2>/dev/null; then echo $d $(cat $d/numa_node) $(cat $d/local_cpulist); fi; done) > $O/r02zz_topology.txt 2>&1
python -m pytest tests -m gpu -q > $O/r02zz_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02zz_pytest.log
tail -3 $O/r02zz_pytest.log
python __graft_entry__.py smoke > $O/r02zz_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r02zz_smoke.log; tail -2 $O/r02zz_smoke.log
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/r02zz_bench_reference.json 2> $O/r02zz_bench_reference.err
python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r02zz_bench_n1.json 2> $O/r02zz_bench_n1.err
python - <<PY
import json
d=json.loads(open('$O/r02zz_bench_n1.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'warmup', d['warmup'], 'floor', round(d['step_floor_frac'],3), 'roofline', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value']), 'cpu', round(d['cpu_baseline']['value'],1))
r=json.loads(open('$O/r02zz_bench_reference.json').read().strip().splitlines()[-1])
print('reference arm', r['value'], r['cpu_baseline']['kind'])
PY
