"""ncu target: the fused head kernel on the spp-608 batch-64 76x76 scale (C_in 256), three launches.
    python profiles/head_ncu_target.py [scale]     scale: 1 = 38x38 / C_in 512, 2 = 76x76 / C_in 256 (default)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pytorch_yolo_b200 import ops  # noqa: E402
from profiles.head_probe import _spp_inputs  # noqa: E402

k = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = "cuda:0"
B, nc = 64, 80
specs, feats, ws, bs = _spp_inputs(B, dev)
rows = sum(s.rows for s in specs)
offs = [0, specs[0].rows, specs[0].rows + specs[1].rows]
buf = ops.Buffers(dev, B, rows, nc)
wp = torch.zeros(256, ws[k].shape[1], device=dev)
wp[:255] = ws[k]
hw = ops.HeadWeights(wp, bs[k].float(), 1.0, 255)
for _ in range(3):
    ops.head_decode_compact([feats[k]], [hw], [specs[k]], [offs[k]], rows, nc, 0.3, buf)
torch.cuda.synchronize()
print("candidates", int(buf.meta[:B].sum()), "overflow", int(buf.meta[B]))
