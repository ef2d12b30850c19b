"""ncu target, end of round 2: the candidate stage and the three NMS kernels (bucket with its work lists, the segment stage with
packed groups, finalize) once per configuration: spp-608 batch 64 at conf 0.3 and 0.001, tiny-416 batch 1024 at conf 0.3."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pytorch_yolo_b200 import ops, synth  # noqa: E402

dev = "cuda:0"
for wl, B, conf in (("spp-608", 64, 0.3), ("spp-608", 64, 0.001), ("tiny-416", 1024, 0.3)):
    w = synth.WORKLOADS[wl]
    heads = synth.synth_heads(wl, B, "B", seed=1234, device=dev)
    specs = [ops.scale_spec(a, g, g, w["img_size"]) for a, g in zip(w["anchors"], w["grids"])]
    buf = ops.Buffers(dev, B, synth.anchors_per_image(wl), w["nc"])
    out, out_row = buf.new_outputs()
    ops.decode_compact(heads, specs, w["nc"], conf, buf)
    ops.nms(buf, 0.5, out, out_row)                     # library default residency (16 warps per SM and list)
    torch.cuda.synchronize()
    del heads, buf, out, out_row
print("done")
