"""Drop-in for the reference's detection head, ``pytorch_yolo.models.yolo_layer.YOLOLayer``
(reference models/yolo_layer.py:25-111), backed by the sm_100a decode kernels.

Same constructor, same attributes (``anchors, n_anchors, n_classes, n_grids, n_x_grids, n_y_grids,
img_size, stride, grid_xy, anchor_vec, anchor_wh, all_anchors, onnx`` -- read by the reference's
``build_targets`` and OpenVINO exporter), same ``forward(p, img_size)`` return values:

* training: the permuted raw tensor ``(B, na, ny, nx, 5+nc)``;
* eval: ``(io, p)`` with ``io`` = decoded ``(B, na*ny*nx, 5+nc)`` and ``p`` the permuted raw tensor.

Two documented deviations: in eval mode ``p`` is returned as a permuted *view* of the input (same
shape and values, no copy -- the reference materialises it with ``.contiguous()``); and the ONNX
export branch (yolo_layer.py:73-88) is out of scope and raises.

``decode_layers`` / ``detect_layers`` are what a model's ``forward`` calls instead of
"each layer, then ``torch.cat(io, 1)``" (reference models/yolov3_spp.py:151-164): all scales in one
launch, straight into the concatenated tensor -- or, fused with the confidence filter and NMS, with
no dense tensor at all.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
from torch import nn

from .. import ops


class YOLOLayer(nn.Module):
    def __init__(self, anchors, nc, all_anchors, onnx=False, in_tensor=None, img_size=None):
        super().__init__()
        self.anchors = torch.Tensor(anchors)
        self.n_anchors = len(anchors)
        self.anchor_wh = 0
        self.n_classes = nc
        self.n_grids = 0
        self.n_x_grids = 0
        self.n_y_grids = 0
        self.img_size = 0
        self.stride = 0
        self.grid_xy = 0
        self.anchor_vec = 0
        self.onnx = onnx
        self.all_anchors = all_anchors
        self._spec: Optional[ops.ScaleSpec] = None
        if onnx and (in_tensor is None or img_size is None):
            raise ValueError('With onnx flag need in_tensor and img_size')      # yolo_layer.py:47-48
        elif onnx:
            _, _, ny, nx = in_tensor.shape
            self.img_size = img_size
            self.n_x_grids = nx
            self.n_y_grids = ny
            self.create_grids(in_tensor.device)

    # reference models/yolo_layer.py:101-111 -- same attributes, same values, same lazy caching
    def create_grids(self, device='cpu'):
        ny, nx = self.n_y_grids, self.n_x_grids
        self.stride = self.img_size / max(nx, ny)
        ys = torch.arange(ny).view(ny, 1).expand(ny, nx)
        xs = torch.arange(nx).view(1, nx).expand(ny, nx)
        self.grid_xy = torch.stack((xs, ys), 2).to(device).float().view(1, 1, ny, nx, 2)
        self.anchor_vec = self.anchors.to(device) / self.stride
        self.anchor_wh = self.anchor_vec.view(1, self.n_anchors, 1, 1, 2)
        self.n_grids = torch.Tensor((nx, ny)).to(device)
        self._spec = ops.scale_spec(self.anchors.tolist(), ny, nx, self.img_size)

    def _prepare(self, p: torch.Tensor, img_size) -> ops.ScaleSpec:
        ny, nx = p.shape[-2], p.shape[-1]
        if not self.onnx and (self.n_x_grids, self.n_y_grids) != (nx, ny):      # yolo_layer.py:59-63
            self.img_size = img_size
            self.n_x_grids = nx
            self.n_y_grids = ny
            self.create_grids(p.device)
        return self._spec

    def _raw_view(self, p: torch.Tensor) -> torch.Tensor:
        bs = p.shape[0]
        return p.view(bs, self.n_anchors, self.n_classes + 5, self.n_y_grids, self.n_x_grids).permute(0, 1, 3, 4, 2)

    def forward(self, p, img_size):
        spec = self._prepare(p, img_size)
        if self.training:
            return self._raw_view(p).contiguous()                                # yolo_layer.py:67-72
        if self.onnx:
            raise NotImplementedError("the ONNX/OpenVINO export branch (reference yolo_layer.py:73-88) is out of "
                                      "scope of pytorch_yolo_b200; export with the reference layer")
        io = ops.decode_dense([p], [spec], self.n_classes)                       # yolo_layer.py:90-99
        return io, self._raw_view(p)


def _specs(layers: Sequence[YOLOLayer], heads: Sequence[torch.Tensor], img_size) -> List[ops.ScaleSpec]:
    if len(layers) != len(heads):
        raise ValueError("one head tensor per YOLO layer is required")
    nc = layers[0].n_classes
    if any(l.n_classes != nc for l in layers):
        raise ValueError("all layers must predict the same number of classes")
    return [l._prepare(h, img_size) for l, h in zip(layers, heads)]


def decode_layers(layers: Sequence[YOLOLayer], heads: Sequence[torch.Tensor], img_size):
    """Eval-mode replacement of ``[yolo_k(branch_k, img_size)]`` + ``torch.cat(io, 1)``
    (reference models/yolov3_spp.py:151-164): returns ``(pred (B, N, 5+nc), tuple_of_p)``."""
    specs = _specs(layers, heads, img_size)
    pred = ops.decode_dense(list(heads), specs, layers[0].n_classes)
    return pred, tuple(l._raw_view(h) for l, h in zip(layers, heads))


def detect_layers(layers: Sequence[YOLOLayer], heads: Sequence[torch.Tensor], img_size,
                  conf_thres: float = 0.5, nms_thres: float = 0.5, return_rows: bool = False):
    """The whole hot path fused: decode + confidence filter + compaction + NMS.  Returns what
    ``non_max_suppression(model(x)[0], conf_thres, nms_thres)`` returns (list of (n,7) | None)."""
    from ..detect import detect
    specs = _specs(layers, heads, img_size)
    return detect(list(heads), specs, layers[0].n_classes, conf_thres, nms_thres, return_rows=return_rows)
