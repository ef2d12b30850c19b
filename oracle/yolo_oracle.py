"""CPU oracle for the YOLO decode + non_max_suppression hot path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  A restatement, in torch
CPU ops of the same granularity the reference uses (so that it can also serve as
the timed ``cpu_baseline`` "port"), of

* ``YOLOLayer.create_grids`` / ``YOLOLayer.forward`` eval branch
  (reference ``pytorch_yolo/models/yolo_layer.py:57-69, 90-99, 101-111``),
* the model-level concat (``pytorch_yolo/models/yolov3_spp.py:163-164``,
  ``pytorch_yolo/models/yolov3_tiny.py:99-100``),
* ``xywh2xyxy`` (``pytorch_yolo/utils/utils.py:46-60``),
* ``bbox_iou`` (``pytorch_yolo/utils/utils.py:63-96``),
* ``non_max_suppression`` with the hard-coded ``'MERGE'`` style
  (``pytorch_yolo/utils/utils.py:200-293``),
* (scope row 8f-3) the head 1x1 convolution that produces the head tensor: ``ConvBlock`` =
  conv + BatchNorm + LeakyReLU (``pytorch_yolo/models/yolo_base.py:19-44``, used at
  ``models/yolov3_spp.py:86,99,111``) or a plain ``nn.Conv2d`` (``models/yolov3_tiny.py:38,42``).

Differences from the reference, all deliberate and documented:

1. *Tie rule.*  The reference orders by ``(-score).argsort()`` which is an unstable
   sort on CPU (SURVEY.md section 0 finding 6).  The oracle uses a *stable* sort, i.e.
   ties in score keep ascending anchor-row order; the final per-image order is
   score descending, ties by (class ascending, in-class emission order).  On
   tie-free inputs the two are identical.
2. *Index tracking.*  Every emitted detection also carries the anchor row it came
   from, so "kept indices" can be compared bit-exactly.
3. ``nms_thres >= 1`` makes the reference loop forever; the
   oracle raises ``ValueError`` instead.

Parity pin: ``tests/golden/*.npz`` produced by the live reference
(``tests/golden/make_golden.py``); checked in ``tests/test_oracle_golden.py``.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

MIN_WH = 2.0          # utils.py:207
MAX_PER_CLASS = 100   # utils.py:247-250


# --------------------------------------------------------------------------- head convolution (8f-3)
def head_conv(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None, negative_slope: float = 1.0,
              bn: Optional[Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, float]] = None) -> torch.Tensor:
    """The producer of a head tensor in fp32 on the CPU: ``conv2d`` 1x1 (weight (n_out, c_in, 1, 1), optional bias), then the
    eval-mode BatchNorm ``bn = (gamma, beta, running_mean, running_var, eps)`` when given, then LeakyReLU -- the op
    sequence of ConvBlock.forward (yolo_base.py:30-44); with bn=None and negative_slope=1 it is the plain nn.Conv2d head
    of yolov3_tiny.py:38,42."""
    y = torch.nn.functional.conv2d(x, weight.view(weight.shape[0], -1, 1, 1), bias)                 # yolo_base.py:31-36
    if bn is not None:
        gamma, beta, mean, var, eps = bn
        y = torch.nn.functional.batch_norm(y, mean, var, gamma, beta, training=False, eps=eps)    # :37
    if negative_slope != 1.0:
        y = torch.nn.functional.leaky_relu(y, negative_slope)                                      # :38
    return y


# --------------------------------------------------------------------------- decode
def scale_constants(anchors, ny: int, nx: int, img_size):
    """stride (python float), anchor_vec fp32 (na,2), grid fp32 (ny,nx,2) -- yolo_layer.py:101-111."""
    stride = img_size / max(nx, ny)                                   # :102 python float
    ys = torch.arange(ny).view(ny, 1).expand(ny, nx)
    xs = torch.arange(nx).view(1, nx).expand(ny, nx)
    grid = torch.stack((xs, ys), 2).float()                            # :105-106 ch0 = x, ch1 = y
    anchor_vec = torch.tensor(anchors, dtype=torch.float32).view(-1, 2) / stride   # :109
    return stride, anchor_vec, grid


def decode_scale(head: torch.Tensor, anchors, n_classes: int, img_size) -> torch.Tensor:
    """Eval-branch decode of one head: (B, na*(5+nc), ny, nx) -> (B, na*ny*nx, 5+nc).

    yolo_layer.py:57-69 (reshape), :90-99 (decode).  Each torch op rounds to fp32
    on its own, exactly as the reference's sequence of in-place slice updates does.
    """
    bs, _, ny, nx = head.shape
    na = len(anchors)
    no = n_classes + 5
    stride, anchor_vec, grid = scale_constants(anchors, ny, nx, img_size)
    p = head.view(bs, na, no, ny, nx).permute(0, 1, 3, 4, 2).contiguous()          # :67-69
    out = torch.empty_like(p)
    out[..., 0:2] = (torch.sigmoid(p[..., 0:2]) + grid.view(1, 1, ny, nx, 2)) * stride     # :91, :94
    out[..., 2:4] = (torch.exp(p[..., 2:4]) * anchor_vec.view(1, na, 1, 1, 2)) * stride    # :92, :94
    out[..., 4:] = torch.sigmoid(p[..., 4:])                                               # :93
    if n_classes == 1:
        out[..., 5] = 1                                                                    # :95-96
    return out.view(bs, -1, no)                                                            # :99


def decode_heads(heads: Sequence[torch.Tensor], anchors_per_scale, n_classes: int, img_size) -> torch.Tensor:
    """All scales in model order, concatenated on the row axis -- yolov3_spp.py:163-164."""
    return torch.cat([decode_scale(h, a, n_classes, img_size) for h, a in zip(heads, anchors_per_scale)], 1)


# --------------------------------------------------------------------------- box helpers
def centre_to_corner(b: torch.Tensor) -> torch.Tensor:
    """(x, y, w, h) -> (x1, y1, x2, y2) -- utils.py:46-60."""
    half_w = b[:, 2] / 2
    half_h = b[:, 3] / 2
    return torch.stack((b[:, 0] - half_w, b[:, 1] - half_h, b[:, 0] + half_w, b[:, 1] + half_h), 1)


def iou_one_to_many(a: torch.Tensor, m: torch.Tensor) -> torch.Tensor:
    """IoU of corner box ``a`` (4,) with each row of ``m`` (n,4) -- utils.py:63-96, x1y1x2y2 branch.

    Operation order is the reference's: inter = clamp(dx,0)*clamp(dy,0);
    union = ((area_a + 1e-16) + area_m) - inter; iou = inter / union.
    """
    m = m.t()
    dx = (torch.min(a[2], m[2]) - torch.max(a[0], m[0])).clamp(0)
    dy = (torch.min(a[3], m[3]) - torch.max(a[1], m[1])).clamp(0)
    inter = dx * dy
    union = ((a[2] - a[0]) * (a[3] - a[1]) + 1e-16) + (m[2] - m[0]) * (m[3] - m[1]) - inter
    return inter / union


# --------------------------------------------------------------------------- candidates
def select_candidates(pred: torch.Tensor, conf_thres: float, write_back: bool = True):
    """Score, class and filter of one image -- utils.py:210-234.

    pred: (N, 5+nc) fp32.  With ``write_back`` the product obj*max_cls is stored into
    ``pred[:, 4]`` like the reference does (utils.py:213, a view into the caller's tensor).
    Returns (rows int64 (K,), det fp32 (K,7) = x1,y1,x2,y2,score,cls_conf,cls) in row order.
    """
    cls_conf, cls_id = pred[:, 5:].max(1)                 # :212 first index on ties
    if write_back:
        pred[:, 4] *= cls_conf                            # :213
        score = pred[:, 4]
    else:
        score = pred[:, 4] * cls_conf
    ok = score > conf_thres                               # :216 strict
    ok = ok & (pred[:, 2:4] > MIN_WH).all(1)              # :217
    if write_back:
        ok = ok & torch.isfinite(pred).all(1)             # :218 (column 4 already holds the product)
    else:
        ok = ok & torch.isfinite(pred[:, :4]).all(1) & torch.isfinite(pred[:, 5:]).all(1) & torch.isfinite(score)
    rows = ok.nonzero().view(-1)
    sel = pred[rows]
    det = torch.cat((centre_to_corner(sel[:, :4]),        # :231
                     score[rows].unsqueeze(1),
                     cls_conf[rows].unsqueeze(1),
                     cls_id[rows].unsqueeze(1).float()), 1)   # :227-234
    return rows, det


def merge_nms_class(det: torch.Tensor, rows: torch.Tensor, nms_thres: float):
    """MERGE-style greedy suppression of one class list already ordered by score -- utils.py:244-275.

    Returns (kept_det (k,7), kept_rows (k,), cluster_sizes list[int]).
    """
    if len(det) == 1:                                     # :244-246 emitted untouched
        return det, rows, [1]
    det = det[:MAX_PER_CLASS].clone()                     # :247-250
    rows = rows[:MAX_PER_CLASS]
    out_det, out_rows, sizes = [], [], []
    while len(det):
        if len(det) == 1:                                 # :268-270 last survivor, unmerged
            out_det.append(det)
            out_rows.append(rows[:1])
            sizes.append(1)
            break
        hit = iou_one_to_many(det[0], det[:, :4]) > nms_thres      # :271 includes itself, strict >
        w = det[hit, 4:5]
        merged = det[:1].clone()
        merged[0, :4] = (w * det[hit, :4]).sum(0) / w.sum()         # :272-273
        out_det.append(merged)
        out_rows.append(rows[:1])
        sizes.append(int(hit.sum()))
        det = det[hit == 0]                                         # :275
        rows = rows[hit == 0]
    return torch.cat(out_det), torch.cat(out_rows), sizes


def nms_image(pred: torch.Tensor, conf_thres: float, nms_thres: float, write_back: bool = True):
    """One image: returns (det (n,7) | None, rows (n,) | None) -- utils.py:210-291."""
    rows, det = select_candidates(pred, conf_thres, write_back)
    if len(rows) == 0:                                    # :223-224
        return None, None
    order = torch.sort(-det[:, 4], stable=True).indices   # :237 with the documented stable tie rule
    det, rows = det[order], rows[order]
    kept_det, kept_rows = [], []
    for c in det[:, -1].unique():                         # :241 ascending class id
        sel = det[:, -1] == c
        d, r, _ = merge_nms_class(det[sel], rows[sel], nms_thres)
        kept_det.append(d)
        kept_rows.append(r)
    kept_det = torch.cat(kept_det)                        # :290
    kept_rows = torch.cat(kept_rows)
    final = torch.sort(-kept_det[:, 4], stable=True).indices      # :291
    return kept_det[final], kept_rows[final]


def non_max_suppression_indexed(prediction, conf_thres: float = 0.5, nms_thres: float = 0.5,
                                write_back: bool = True
                                ) -> Tuple[List[Optional[torch.Tensor]], List[Optional[torch.Tensor]]]:
    """Reference-shaped call that also returns the kept anchor rows per image."""
    if not nms_thres < 1:
        raise ValueError("nms_thres >= 1 never terminates in the reference (utils.py:266-275)")
    dets: List[Optional[torch.Tensor]] = [None] * len(prediction)
    idxs: List[Optional[torch.Tensor]] = [None] * len(prediction)
    for i, pred in enumerate(prediction):                 # :210
        dets[i], idxs[i] = nms_image(pred, conf_thres, nms_thres, write_back)
    return dets, idxs


def non_max_suppression(prediction, conf_thres: float = 0.5, nms_thres: float = 0.5):
    """Same signature and return value as the reference function (utils.py:200)."""
    return non_max_suppression_indexed(prediction, conf_thres, nms_thres)[0]


def detect(heads, anchors_per_scale, n_classes, img_size, conf_thres, nms_thres):
    """decode + NMS in one call: what ``model(x)`` followed by ``non_max_suppression`` computes."""
    pred = decode_heads(heads, anchors_per_scale, n_classes, img_size)
    return non_max_suppression_indexed(pred, conf_thres, nms_thres)


# --------------------------------------------------------------------------- post-NMS epilogue
def scale_coords(img1_shape, coords: torch.Tensor, img0_shape) -> torch.Tensor:
    """Letterbox un-padding of xyxy boxes, in place -- utils.py:296-303."""
    gain = max(img1_shape) / max(img0_shape)                                   # :298
    coords[:, [0, 2]] -= (img1_shape[1] - img0_shape[1] * gain) / 2            # :299
    coords[:, [1, 3]] -= (img1_shape[0] - img0_shape[0] * gain) / 2            # :300
    coords[:, :4] /= gain                                                      # :301
    coords[:, :4] = coords[:, :4].clamp(min=0)                                 # :302
    return coords


def records_from_results(data: dict, targets, imgs_path, orig_shapes, cur_shape) -> dict:
    """Per-image detection records -- utils.py:306-327."""
    for i, pred in enumerate(targets):
        if pred is None:
            continue
        pred[:, :4] = scale_coords(cur_shape, pred[:, :4], orig_shapes[i]).round()        # :313
        for x1, y1, x2, y2, conf, _cls_conf, cls in pred.detach().cpu().numpy():          # :314
            data.setdefault(imgs_path[i], []).append({'type': int(cls), 'score': float(conf), 'left': int(x1),
                                                      'top': int(y1), 'right': int(x2), 'bottom': int(y2)})
    return data
