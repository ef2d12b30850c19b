"""Probe: how much of the NMS stage hides behind the decode kernel in the pipelined detector.
Times, for spp-608 batch 64 conf 0.3 and several depths: decode only, NMS only (on fixed candidates), both."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pytorch_yolo_b200 import ops, synth                       # noqa: E402
from pytorch_yolo_b200.detect import Detector, PipelinedDetector  # noqa: E402

dev = torch.device("cuda:0")
wl, B, conf = "spp-608", 64, 0.3
w = synth.WORKLOADS[wl]
heads = synth.synth_heads(wl, B, "B", seed=1234, device=dev)
specs = [ops.scale_spec(a, g, g, w["img_size"]) for a, g in zip(w["anchors"], w["grids"])]


def variant(mode):
    def enqueue(self, hs):
        if mode in ("both", "decode"):
            ops.decode_compact(hs, self.specs, self.nc, self.conf_thres, self.buf, variant=self.variant)
        if mode in ("both", "nms"):
            ops.nms(self.buf, self.nms_thres, self.out, self.out_row, out_ptrs=self.out_ptrs)
        self.buf.meta_host.copy_(self.buf.meta, non_blocking=True)
    return enqueue


def run(mode, depth, steps=1500):
    pipe = PipelinedDetector(specs, w["nc"], B, dev, conf, 0.5, depth=depth)
    for d in pipe.lanes:                       # candidates must exist for the NMS-only variant
        ops.decode_compact(heads, specs, w["nc"], conf, d.buf)
    torch.cuda.synchronize()
    Detector._enqueue = variant(mode)
    pend = []
    def loop(k):
        for _ in range(k):
            pend.append(pipe.submit(heads))
            if len(pend) >= depth:
                pipe.lanes[pend.pop(0) % depth]._stream.synchronize()
        while pend:
            pipe.lanes[pend.pop(0) % depth]._stream.synchronize()
    loop(30)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    loop(steps)
    pipe.drain()
    e1.record()
    torch.cuda.synchronize()
    print(f"{mode:7s} depth {depth}: {e0.elapsed_time(e1) / steps * 1000:7.1f} us/step", flush=True)


orig = Detector._enqueue
for depth in (1, 4, 6):
    for mode in ("decode", "nms", "both"):
        run(mode, depth)
Detector._enqueue = orig
