"""GPU: the persistent Detector (CUDA-graph replay), the pipelined detector and the host-buffer path must give
exactly what the one-shot fused call gives."""
import pytest
import torch

from pytorch_yolo_b200 import YOLOLayer, detect_layers, ops, synth
from pytorch_yolo_b200.detect import Detector, PipelinedDetector

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _setup(workload, batch, kind, seed):
    w = synth.WORKLOADS[workload]
    layers = [YOLOLayer(a, w["nc"], w["anchors"]).eval() for a in w["anchors"]]
    heads = [h.to(DEV) for h in synth.synth_heads(workload, batch, kind, seed=seed)]
    specs = [ops.scale_spec(a, g, g, w["img_size"]) for a, g in zip(w["anchors"], w["grids"])]
    return w, layers, heads, specs


def _same(a, b):
    assert len(a) == len(b)
    for x, y in zip(a, b):
        assert (x is None) == (y is None)
        if x is not None:
            assert torch.equal(x, y)


@pytest.mark.parametrize("use_graph", [False, True])
def test_detector_matches_one_shot(use_graph):
    w, layers, heads, specs = _setup("tiny-416", 6, "B", 81)
    want, want_rows = detect_layers(layers, heads, 416, 0.3, 0.5, return_rows=True)
    det = Detector(specs, w["nc"], 6, DEV, 0.3, 0.5, use_graph=use_graph)
    for _ in range(3):                                   # replay must be idempotent
        got, rows = det.run(heads, return_rows=True, clone=True)
        _same(got, want)
        _same(rows, want_rows)
    # new data in the same (static) tensors
    fresh = synth.synth_heads("tiny-416", 6, "B", seed=82)
    for h, f in zip(heads, fresh):
        h.copy_(f)
    _same(det.run(heads, clone=True), detect_layers(layers, heads, 416, 0.3, 0.5))


def test_pipelined_detector_keeps_batches_apart():
    w, layers, heads_a, specs = _setup("mini-160", 4, "B", 83)
    heads_b = [h.to(DEV) for h in synth.synth_heads("mini-160", 4, "B", seed=84)]
    want_a = detect_layers(layers, heads_a, 160, 0.05, 0.5)
    want_b = detect_layers(layers, heads_b, 160, 0.05, 0.5)
    pipe = PipelinedDetector(specs, w["nc"], 4, DEV, 0.05, 0.5, depth=2)
    for _ in range(3):
        ta = pipe.submit(heads_a)
        tb = pipe.submit(heads_b)
        _same(pipe.collect(ta, clone=True), want_a)
        _same(pipe.collect(tb, clone=True), want_b)


def test_run_from_host_buffers():
    w, layers, heads, specs = _setup("tiny-416", 3, "B", 85)
    want = detect_layers(layers, heads, 416, 0.3, 0.5)
    det = Detector(specs, w["nc"], 3, DEV, 0.3, 0.5)
    host_heads = [h.cpu().pin_memory() for h in heads]
    staging = [torch.empty_like(h) for h in heads]
    host_out = torch.empty(3, det.buf.out_cap, 7).pin_memory()
    kept, out, h2d, d2h = det.run_from_host(host_heads, staging, host_out)
    assert h2d == sum(h.numel() * 4 for h in heads) and d2h > 0
    for i, n in enumerate(kept.tolist()):
        if want[i] is None:
            assert n == 0
        else:
            assert torch.equal(out[i, :n], want[i].cpu())


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_sharded_detector_matches_single_gpu():
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    n = min(torch.cuda.device_count(), 4)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", "29533",
                        os.path.join(root, "tests", "multi_gpu_check.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "multi-gpu check: ok" in r.stdout


def test_eval_pipeline_matches_sequential_drop_ins():
    """the overlapped evaluation loop must produce exactly the records the sequential
    reference-shaped calls produce (model forward -> non_max_suppression -> _dict_from_results)."""
    from torch import nn
    from pytorch_yolo_b200 import decode_layers, non_max_suppression
    from pytorch_yolo_b200.pipeline import EvalPipeline
    from pytorch_yolo_b200.utils.utils import dict_from_results

    anchors = synth.TINY_ANCHORS

    class TinyHeadModel(nn.Module):                              # stands in for the (out of scope) backbone
        def __init__(self):
            super().__init__()
            self.c1 = nn.Conv2d(3, 255, 16, stride=16)
            self.c2 = nn.Conv2d(3, 255, 32, stride=32)
            self.yolo1 = YOLOLayer(anchors[0], 80, anchors)
            self.yolo2 = YOLOLayer(anchors[1], 80, anchors)

        @property
        def yolo_layers(self):
            return self.yolo1, self.yolo2

        def _forward_encoder(self, x):
            return self.c1(x) * 3.0, self.c2(x) * 3.0

    torch.manual_seed(5)
    model = TinyHeadModel().to(DEV).eval()
    g = torch.Generator().manual_seed(6)
    batches = []
    for k in range(5):
        b = 3 if k < 4 else 2                                    # a ragged last batch -> a second detector shape
        imgs = torch.rand(b, 3, 128, 160, generator=g)
        paths = [f"b{k}_i{i}.jpg" for i in range(b)]
        shapes = [(int(100 + 50 * i + 7 * k), int(200 + 31 * i)) for i in range(b)]
        batches.append((imgs, paths, shapes))

    got = EvalPipeline(model, DEV, conf_thresh=0.3, nms_thresh=0.45, depth=2).run(batches)

    want = {}
    with torch.no_grad():
        for imgs, paths, shapes in batches:
            x = imgs.to(DEV)
            pred, _ = decode_layers(model.yolo_layers, model._forward_encoder(x), max(x.shape[-2:]))
            dets = non_max_suppression(pred, 0.3, 0.45)
            dict_from_results(want, dets, paths, shapes, tuple(x.shape[-2:]))
    assert sum(len(v) for v in want.values()) > 10
    assert got == want


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_second_device_and_side_stream_in_one_process():
    """The wrappers follow the tensors' device and the caller's current stream (no hidden use of cuda:0 / stream 0)."""
    w, layers, heads0, specs = _setup("tiny-416", 4, "B", 91)
    want = [None if d is None else d.cpu() for d in detect_layers(layers, heads0, 416, 0.3, 0.5)]
    heads1 = [h.to("cuda:1") for h in heads0]
    layers1 = [YOLOLayer(a, w["nc"], w["anchors"]).eval() for a in w["anchors"]]
    side = torch.cuda.Stream("cuda:1")
    with torch.cuda.stream(side):
        got = detect_layers(layers1, heads1, 416, 0.3, 0.5)
    side.synchronize()
    for g, x in zip(got, want):
        assert (g is None) == (x is None)
        if g is not None:
            assert g.device == torch.device("cuda:1") and torch.equal(g.cpu(), x)
