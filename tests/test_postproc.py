"""Post-NMS epilogue: scale_coords + _dict_from_results (reference utils/utils.py:296-327).
CPU: the oracle restatement against the golden vectors of the live reference.  GPU: the kernels against both."""
import json

import numpy as np
import pytest
import torch

from oracle import yolo_oracle
from tests.helpers import load_golden


def _case():
    g = load_golden("postproc")
    n = int(g["n_images"])
    dets = [torch.from_numpy(g[f"det{i}"].copy()) if f"det{i}" in g else None for i in range(n)]
    scaled = [torch.from_numpy(g[f"scaled{i}"]) if f"scaled{i}" in g else None for i in range(n)]
    shapes = [tuple(int(v) for v in s) for s in g["orig_shapes"]]
    cur = tuple(int(v) for v in g["cur_shape"])
    paths = [f"img_{i if i < 2 else i - 1}.jpg" if i != 2 else "none.jpg" for i in range(n)]
    return dets, scaled, shapes, cur, paths, json.loads(str(g["records"]))


def test_oracle_scale_coords_and_records_match_golden():
    dets, scaled, shapes, cur, paths, records = _case()
    for d, s, o in zip(dets, scaled, shapes):
        if d is not None:
            got = yolo_oracle.scale_coords(cur, d[:, :4].clone(), o)
            assert torch.equal(got, s)
    got = yolo_oracle.records_from_results({}, [None if d is None else d.clone() for d in dets], paths, shapes, cur)
    assert got == records


@pytest.mark.gpu
def test_gpu_scale_coords_and_records_bit_exact():
    from pytorch_yolo_b200.utils.utils import dict_from_results, scale_coords
    dets, scaled, shapes, cur, paths, records = _case()
    for d, s, o in zip(dets, scaled, shapes):
        if d is None:
            continue
        rows = d.clone().to("cuda:0")
        view = rows[:, :4]                                    # a strided view of the 7-column rows, as in utils.py:313
        out = scale_coords(cur, view, o)
        assert out.data_ptr() == view.data_ptr()              # in place, returns its argument like the reference
        assert torch.equal(rows[:, :4].cpu(), s) and torch.equal(rows[:, 4:].cpu(), d[:, 4:])
        tight = d[:, :4].clone().to("cuda:0")                 # contiguous (n,4) input
        assert torch.equal(scale_coords(cur, tight, o).cpu(), s)
    got = dict_from_results({}, [None if d is None else d.clone().to("cuda:0") for d in dets], paths, shapes, cur)
    assert got == records


@pytest.mark.gpu
def test_detector_batch_epilogue_matches_reference_semantics():
    from pytorch_yolo_b200 import ops, synth
    from pytorch_yolo_b200.detect import Detector
    wl, B = "tiny-416", 4
    w = synth.WORKLOADS[wl]
    specs = [ops.scale_spec(a, g, g, w["img_size"]) for a, g in zip(w["anchors"], w["grids"])]
    heads = [h.to("cuda:0") for h in synth.synth_heads(wl, B, "B", seed=17)]
    det = Detector(specs, w["nc"], B, "cuda:0", 0.3, 0.5, use_graph=False)
    before = [None if d is None else d.cpu().clone() for d in det.run(heads)]
    shapes = [(480, 640), (1080, 1920), (416, 416), (333, 500)]
    det.scale_to_original((416, 416), shapes)
    torch.cuda.synchronize()
    _, kept = det.counts()
    after = ops.ragged(det.out, det.out_row, kept)
    for b, a, o in zip(before, after, shapes):
        if b is None:
            assert a is None
            continue
        want = b.clone()
        want[:, :4] = yolo_oracle.scale_coords((416, 416), want[:, :4], o).round()
        assert torch.equal(a.cpu(), want)
    # values at exactly .5 round half to even, negatives clamp to 0
    t = torch.tensor([[0.5, 1.5, 2.5, -3.0]], device="cuda:0")
    from pytorch_yolo_b200.utils.utils import _scale_rows
    assert _scale_rows(t, (10, 10), (10, 10), True).cpu().tolist() == [[0.0, 2.0, 2.0, 0.0]]
