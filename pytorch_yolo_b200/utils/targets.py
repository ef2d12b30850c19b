"""Training-side consumer of the YOLOLayer constants: ``build_targets``.

Drop-in for the reference's ``build_targets(model, targets)`` (utils/utils.py:160-197): same arguments, same return value
``(txy, twh, tcls, indices)`` -- per YOLO layer the regression targets and the ``(image, anchor, grid y, grid x)`` index
tuple ``compute_loss`` (utils.py:124-157) consumes -- computed for all layers by one kernel (``csrc/targets.cu``) instead of
about thirty small torch launches per layer.  ``model`` needs what the reference reads: ``hyper_params['iou_thresh']`` and
``yolo_layers`` whose ``n_x_grids / n_y_grids / anchor_vec / n_classes`` are populated (``YOLOLayer.create_grids``: after
one forward pass, exactly as in the reference).  ``targets``: (nt, 6) fp32 CUDA tensor ``[image, class, x, y, w, h]``.
"""
from __future__ import annotations

from typing import List, Tuple

import torch

from .. import _lib
from ..ops import _require_cuda, _stream_ptr


def build_targets(model, targets: torch.Tensor):
    _require_cuda(targets, "targets")
    if targets.dim() != 2 or targets.shape[1] != 6:
        raise ValueError("targets must be (nt, 6): image, class, x, y, w, h")
    lib = _lib.load()
    layers = list(model.yolo_layers)
    if not 1 <= len(layers) <= _lib.MAX_SCALES:
        raise ValueError(f"1..{_lib.MAX_SCALES} YOLO layers are supported")
    iou_thres = float(model.hyper_params['iou_thresh'])                    # utils.py:162
    t = targets if targets.is_contiguous() else targets.contiguous()
    nt, dev = t.shape[0], t.device
    arr = (_lib.TargetLayer * len(layers))()
    outs = []
    n = max(nt, 1)
    for k, layer in enumerate(layers):
        if not layer.n_x_grids or not torch.is_tensor(layer.anchor_vec):
            raise ValueError("run one forward pass first: the layer's grid constants are created lazily (yolo_layer.py:59-63)")
        idx = torch.empty(5, n, dtype=torch.int64, device=dev)             # b, a, gj, gi, tcls
        reg = torch.empty(2, n, 2, dtype=torch.float32, device=dev)        # txy, twh
        outs.append((idx, reg))
        e = arr[k]
        e.nx, e.ny, e.na = int(layer.n_x_grids), int(layer.n_y_grids), int(layer.n_anchors)
        for a, (aw, ah) in enumerate(layer.anchor_vec.tolist()):
            e.anchor_vec[a][0], e.anchor_vec[a][1] = aw, ah
        e.b, e.a, e.gj, e.gi, e.tcls = (idx[i].data_ptr() for i in range(5))
        e.txy, e.twh = reg[0].data_ptr(), reg[1].data_ptr()
    count = torch.zeros(len(layers), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.yolo_b200_build_targets(t.data_ptr(), nt, arr, len(layers), iou_thres, count.data_ptr(),
                                               _stream_ptr(dev)), "yolo_b200_build_targets")
    kept = count.tolist()                                                  # the one host sync: output lengths
    txy: List[torch.Tensor] = []
    twh: List[torch.Tensor] = []
    tcls: List[torch.Tensor] = []
    indices: List[Tuple[torch.Tensor, ...]] = []
    for (idx, reg), m, layer in zip(outs, kept, layers):
        indices.append((idx[0, :m], idx[1, :m], idx[2, :m], idx[3, :m]))   # utils.py:185
        txy.append(reg[0, :m])
        twh.append(reg[1, :m])
        tcls.append(idx[4, :m])
        if m and int(idx[4, :m].max()) > layer.n_classes:                  # utils.py:195-196
            raise AssertionError('Target classes exceed model classes')
    return txy, twh, tcls, indices
