#!/bin/bash
# round 2, pass i: TMA dense decode (parity + timing), eval-loop golden tests, NMS with residency option
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r02i_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02i_pytest.log
tail -5 $O/r02i_pytest.log
python - > $O/r02i_dense.txt 2>&1 <<'PY'
import json, torch, sys
sys.path.insert(0, '.')
from pytorch_yolo_b200 import ops, synth
peak = json.load(open('MEASURED_PEAKS.json'))['hbm_gbs']
for wl, B in (("spp-608", 64), ("tiny-416", 1024), ("spp-1024", 64)):
    w = synth.WORKLOADS[wl]
    heads = synth.synth_heads(wl, B, "A", seed=1, device="cuda:0")
    specs = [ops.scale_spec(a, g, g, w["img_size"]) for a, g in zip(w["anchors"], w["grids"])]
    n = synth.anchors_per_image(wl)
    out = torch.empty(B, n, 85, device="cuda:0")
    T = 2 * B * synth.head_bytes_per_image(wl)
    for variant in ("ldg", "tma"):
        for _ in range(5): ops.decode_dense(heads, specs, 80, out=out, variant=variant)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(30): ops.decode_dense(heads, specs, 80, out=out, variant=variant)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 30 * 1e3
        print(f"{wl} b{B} decode_dense {variant}: {us:.1f} us  {T/us/1e3:.0f} GB/s  {T/us/1e3/peak:.3f} of measured peak")
    del heads, out
PY
cat $O/r02i_dense.txt
