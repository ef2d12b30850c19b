"""ncu target of round 2: every kernel of the hot path and of the "next" rows once or twice, on the BASELINE cfg-2 shapes
(spp-608 batch 64, SYNTH-B, conf 0.3), plus the NMS kernels at conf 0.001.
    ncu --set full --clock-control none --import-source on -k regex:... python profiles/ncu_target_r02.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pytorch_yolo_b200 import ops, synth  # noqa: E402
from pytorch_yolo_b200.head import HeadDetector  # noqa: E402

dev, wl, B = "cuda:0", "spp-608", 64
w = synth.WORKLOADS[wl]
n = synth.anchors_per_image(wl)
heads = synth.synth_heads(wl, B, "B", seed=1234, device=dev)
specs = [ops.scale_spec(a, g, g, w["img_size"]) for a, g in zip(w["anchors"], w["grids"])]
buf = ops.Buffers(dev, B, n, w["nc"])
out, out_row = buf.new_outputs()
pred = torch.empty(B, n, w["nc"] + 5, device=dev)
for rep in range(2):
    for conf in (0.3, 0.001):
        ops.decode_compact(heads, specs, w["nc"], conf, buf)
        ops.nms(buf, 0.5, out, out_row, seg_warps_per_sm=32)
    ops.decode_dense(heads, specs, w["nc"], out=pred)
    ops.compact_from_dense(pred, 0.3, buf, write_back=False)
del pred
# the segment stage on many small images (packed groups): tiny-416 batch 1024
wt = synth.WORKLOADS["tiny-416"]
heads_t = synth.synth_heads("tiny-416", 1024, "B", seed=1234, device=dev)
specs_t = [ops.scale_spec(a, g, g, wt["img_size"]) for a, g in zip(wt["anchors"], wt["grids"])]
buf_t = ops.Buffers(dev, 1024, synth.anchors_per_image("tiny-416"), wt["nc"])
out_t, row_t = buf_t.new_outputs()
for rep in range(2):
    ops.decode_compact(heads_t, specs_t, wt["nc"], 0.3, buf_t)
    ops.nms(buf_t, 0.5, out_t, row_t)
del heads_t, buf_t, out_t, row_t
feats, convs = synth.synth_head_convs(wl, B, device=dev)
for precision in ("tf32", "fp32x3"):
    det = HeadDetector(convs, specs, w["nc"], B, dev, 0.3, 0.5, precision=precision)
    for rep in range(2):
        det._produce(feats)
torch.cuda.synchronize()
print("done")
