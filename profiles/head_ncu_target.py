"""ncu target: the fused head path of the BASELINE spp-608 batch-64 shapes -- plane padding of the 19x19 map, then ONE
launch of head_decode_compact_kernel over all three scales -- three repetitions.
    python profiles/head_ncu_target.py            all three scales (pad_planes_kernel + head_decode_compact_kernel)
    python profiles/head_ncu_target.py 2          one scale: 0 = 19x19 / C_in 1024, 1 = 38x38 / 512, 2 = 76x76 / 256"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pytorch_yolo_b200 import ops  # noqa: E402
from profiles.head_probe import _spp_inputs  # noqa: E402

only = int(sys.argv[1]) if len(sys.argv) > 1 else None
dev = "cuda:0"
B, nc = 64, 80
specs, feats, ws, bs = _spp_inputs(B, dev)
rows = sum(s.rows for s in specs)
offs = [0, specs[0].rows, specs[0].rows + specs[1].rows]
buf = ops.Buffers(dev, B, rows, nc)
hws = []
for k in range(3):
    wp = torch.zeros(256, ws[k].shape[1], device=dev)
    wp[:255] = ws[k]
    hws.append(ops.HeadWeights(wp, bs[k].float(), 1.0, 255))
xp = torch.zeros(B, 1024, 364, device=dev)
sel = [0, 1, 2] if only is None else [only]
for _ in range(3):
    xs = [ops.pad_feature(feats[0], out=xp) if k == 0 else feats[k] for k in sel]
    ops.head_decode_compact(xs, [hws[k] for k in sel], [specs[k] for k in sel], [offs[k] for k in sel], rows, nc, 0.3, buf)
torch.cuda.synchronize()
print("scales", sel, "candidates", int(buf.meta[:B].sum()), "overflow", int(buf.meta[B]))
