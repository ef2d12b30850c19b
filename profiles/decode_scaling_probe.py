"""Probe: decode_compact time per scale and versus batch size (fit t = t0 + bytes / BW)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pytorch_yolo_b200 import ops, synth  # noqa: E402

dev = "cuda:0"


def timeit(fn, reps=40, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps * 1e3)
    return best


def run(wl, B, scales=None, conf=0.3):
    w = synth.WORKLOADS[wl]
    heads = synth.synth_heads(wl, B, "B", seed=1234, device=dev)
    specs = [ops.scale_spec(a, g, g, w["img_size"]) for a, g in zip(w["anchors"], w["grids"])]
    if scales is not None:
        heads = [heads[k] for k in scales]
        specs = [specs[k] for k in scales]
    n = sum(s.rows for s in specs)
    buf = ops.Buffers(dev, B, n, w["nc"])
    t = timeit(lambda: ops.decode_compact(heads, specs, w["nc"], conf, buf, variant="ldg"))
    nbytes = sum(h.numel() * 4 for h in heads)
    print(f"{wl:9s} B={B:5d} scales={scales} {nbytes / 1e6:8.1f} MB {t:8.1f} us {nbytes / t / 1e3:8.1f} GB/s", flush=True)


for sc in ([0], [1], [2], [1, 2], None):
    run("spp-608", 64, sc)
for sc in ([0], [1], None):
    run("tiny-416", 1024, sc)
for B in (16, 32, 64, 128, 256, 512):
    run("spp-608", B)
for B in (16, 64, 256):
    run("spp-1024", B)
run("spp-608", 64, None, conf=2.0)      # nothing passes: pure streaming, no candidate writes
