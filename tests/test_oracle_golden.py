"""CPU: the oracle restatement against the golden vectors the live reference produced,
and (when /root/reference is present) against the live reference itself."""
import numpy as np
import pytest
import torch

from oracle import ref_loader, yolo_oracle
from pytorch_yolo_b200 import synth
from tests.helpers import DECODE_GOLDEN, NMS_GOLDEN, assert_dets_equal, load_golden, unpack


@pytest.mark.parametrize("name", NMS_GOLDEN)
def test_oracle_nms_matches_golden_bit_exact(name):
    g = load_golden(name)
    pred = torch.from_numpy(g["pred"].copy())
    dets, idxs = yolo_oracle.non_max_suppression_indexed(pred, float(g["conf"]), float(g["nms"]))
    assert_dets_equal(dets, unpack(g["counts"], g["dets"]), what=name)
    # the in-place side effect on column 4 (reference utils.py:213); NaNs compare equal
    np.testing.assert_array_equal(pred[..., 4].numpy(), g["col4_after"])
    for d, i in zip(dets, idxs):
        assert (d is None) == (i is None)
        if d is not None:
            assert len(d) == len(i) and i.dtype == torch.int64


@pytest.mark.parametrize("name", DECODE_GOLDEN)
def test_oracle_decode_and_nms_match_golden(name):
    g = load_golden(name)
    w = synth.WORKLOADS[str(g["workload"])]
    heads = [torch.from_numpy(g[f"head{k}"].copy()) for k in range(len(w["grids"]))]
    pred = yolo_oracle.decode_heads(heads, w["anchors"], w["nc"], w["img_size"])
    assert torch.equal(pred, torch.from_numpy(g["decoded"])), "oracle decode is not bit-identical to the reference"
    dets, _ = yolo_oracle.non_max_suppression_indexed(pred, float(g["conf"]), float(g["nms"]))
    assert_dets_equal(dets, unpack(g["counts"], g["dets"]), what=name)


def test_oracle_tiny416_randinit_golden():
    """BASELINE config 1: YOLOv3-tiny 416x416 batch 1, random init (heads from the reference's encoder)."""
    g = load_golden("tiny416_randinit")
    w = synth.WORKLOADS["tiny-416"]
    heads = [torch.from_numpy(g["head0"].copy()), torch.from_numpy(g["head1"].copy())]
    pred = yolo_oracle.decode_heads(heads, w["anchors"], w["nc"], w["img_size"])
    assert pred.shape == (1, 2535, 85)
    assert torch.equal(pred[:, ::7], torch.from_numpy(g["decoded_every7"]))
    dets, _ = yolo_oracle.non_max_suppression_indexed(pred, float(g["conf"]), float(g["nms"]))
    assert_dets_equal(dets, unpack(g["counts"], g["dets"]), what="tiny416")
    np.testing.assert_array_equal(pred[..., 4].numpy(), g["col4_after"])


def test_oracle_rejects_nms_thres_one():
    with pytest.raises(ValueError):
        yolo_oracle.non_max_suppression(torch.zeros(1, 4, 7), 0.1, 1.0)


def test_oracle_accepts_list_input_and_write_back_flag():
    pred = synth.synth_prediction(2, 200, nc=5, seed=3)
    a, ia = yolo_oracle.non_max_suppression_indexed(pred.clone(), 0.2, 0.5)
    b, ib = yolo_oracle.non_max_suppression_indexed([p.clone() for p in pred], 0.2, 0.5)
    keep = pred.clone()
    c, ic = yolo_oracle.non_max_suppression_indexed(keep, 0.2, 0.5, write_back=False)
    assert torch.equal(keep, pred)
    for x, y, z, i, j, k in zip(a, b, c, ia, ib, ic):
        assert torch.equal(x, y) and torch.equal(x, z) and torch.equal(i, j) and torch.equal(i, k)


@pytest.mark.skipif(not ref_loader.available(), reason="live reference not present (GPU box)")
@pytest.mark.parametrize("seed,ties", [(1, 0), (2, 0), (3, 6), (4, 3)])
def test_oracle_matches_live_reference(seed, ties):
    ref = ref_loader.load()
    pred = synth.synth_prediction(2, 400, nc=12, seed=seed, tie_levels=ties)
    a = pred.clone()
    with ref_loader.stable_argsort():
        want = ref.non_max_suppression(a, 0.15, 0.5)
    b = pred.clone()
    got = yolo_oracle.non_max_suppression(b, 0.15, 0.5)
    assert_dets_equal(got, want, what=f"seed {seed}")
    assert torch.equal(a[..., 4], b[..., 4])
