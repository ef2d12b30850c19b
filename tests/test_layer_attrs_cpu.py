"""Host-side contract of the YOLOLayer drop-in (no GPU): the attributes the reference's build_targets and OpenVINO
exporter read (reference utils/utils.py:166-195, openvino_converter/layers.py:186-194) and the training-mode forward
(models/yolo_layer.py:57-72) are produced by plain torch code and must equal the live reference bit for bit."""
import pytest
import torch

from oracle import ref_loader
from pytorch_yolo_b200 import YOLOLayer
from pytorch_yolo_b200.ops import YoloB200Error

ANCHORS = [(10.0, 13.0), (16.0, 30.0), (33.0, 23.0)]
ALL = (((116, 90), (156, 198), (373, 326)), ((30, 61), (62, 45), (59, 119)), tuple(ANCHORS))


def test_constructor_contract_without_reference():
    l = YOLOLayer(ANCHORS, 80, ALL)
    assert l.n_anchors == 3 and l.n_classes == 80 and l.onnx is False and l.all_anchors == ALL
    assert list(l.parameters()) == [] and list(l.buffers()) == []            # constants are plain attributes (yolo_layer.py:30)
    with pytest.raises(ValueError):
        YOLOLayer(ANCHORS, 80, ALL, onnx=True)                                # yolo_layer.py:47-48
    l.eval()
    with pytest.raises(YoloB200Error):                                        # no CPU path for the decode itself
        l(torch.zeros(1, 255, 4, 4), 128)


@pytest.mark.skipif(not ref_loader.available(), reason="reference checkout not present")
@pytest.mark.parametrize("ny,nx,img", [(13, 13, 416), (19, 19, 608), (10, 16, 512), (76, 76, 608)])
def test_grids_and_training_forward_equal_the_reference(ny, nx, img):
    ref = ref_loader.load()
    ours, theirs = YOLOLayer(ANCHORS, 80, ALL).train(), ref.YOLOLayer(ANCHORS, 80, ALL).train()
    p = torch.randn(2, 255, ny, nx, generator=torch.Generator().manual_seed(ny * 100 + nx))
    a, b = ours(p, img), theirs(p.clone(), img)
    assert a.shape == b.shape == (2, 3, ny, nx, 85) and torch.equal(a, b) and a.is_contiguous()
    assert ours.stride == theirs.stride and ours.img_size == theirs.img_size
    assert (ours.n_x_grids, ours.n_y_grids) == (theirs.n_x_grids, theirs.n_y_grids)
    for name in ("grid_xy", "anchor_vec", "anchor_wh", "n_grids", "anchors"):
        x, y = getattr(ours, name), getattr(theirs, name)
        assert x.shape == y.shape and x.dtype == y.dtype and torch.equal(x, y), name
