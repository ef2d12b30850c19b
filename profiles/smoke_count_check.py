"""Why can smoke()'s detection count move by one between builds?  NMS of the GPU-decoded tensor against the oracle's NMS of the
SAME tensor (bit-exact expected), and the candidates whose score lies within 1e-6 of the threshold."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from oracle import yolo_oracle  # noqa: E402
from pytorch_yolo_b200 import YOLOLayer, decode_layers, detect_layers, non_max_suppression, synth  # noqa: E402

wl = "tiny-416"
w = synth.WORKLOADS[wl]
heads = synth.synth_heads(wl, 2, "B", seed=5)
layers = [YOLOLayer(a, w["nc"], w["anchors"]).eval() for a in w["anchors"]]
dev_heads = [h.to("cuda:0") for h in heads]
pred, _ = decode_layers(layers, dev_heads, w["img_size"])
pred_cpu = pred.cpu().clone()
got, rows = non_max_suppression(pred.clone(), 0.3, 0.5, return_rows=True)
want, wrows = yolo_oracle.non_max_suppression_indexed(pred_cpu.clone(), 0.3, 0.5)
print("gpu nms of gpu-decoded:", [len(g) for g in got], " oracle nms of the same tensor:", [len(x) for x in want])
for g, r, o, orow in zip(got, rows, want, wrows):
    same = g.shape == o.shape and torch.equal(g[:, 4:].cpu(), o[:, 4:]) and torch.equal(r.cpu().long(), orow)
    print("  bit-exact scores / classes / rows:", same)
fused = detect_layers(layers, dev_heads, w["img_size"], 0.3, 0.5)
print("fused:", [len(f) for f in fused])
cpu_pred = yolo_oracle.decode_heads(heads, w["anchors"], w["nc"], w["img_size"])
want2, _ = yolo_oracle.non_max_suppression_indexed(cpu_pred.clone(), 0.3, 0.5)
print("oracle decode + nms:", [len(x) for x in want2])
