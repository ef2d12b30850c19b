"""Drop-in for ``pytorch_yolo.utils.utils.non_max_suppression`` (reference utils/utils.py:200-293).

Same signature, same return value (python list, one entry per image: ``None`` or an fp32 ``(n, 7)``
tensor ``(x1, y1, x2, y2, obj*cls_conf, cls_conf, cls)`` on the prediction's device, ordered by
column 4 descending), same side effect (``prediction[..., 4]`` is multiplied in place by the max class
confidence, utils.py:213).  The work is done by ``compact_from_dense`` + the segmented NMS kernels.

Documented behaviour where the reference is undefined:
* score ties keep ascending anchor-row order (the reference's argsort is unstable);
* ``nms_thres >= 1`` raises ``ValueError`` (the reference loops forever);
* tensors must be float32 on a CUDA device -- there is no CPU path.
"""
from __future__ import annotations

import torch

from .. import ops


def _nms_tensor(pred: torch.Tensor, conf_thres: float, nms_thres: float, return_rows: bool, cap=None):
    ops._require_cuda(pred, "prediction")
    if pred.dim() != 3:
        raise ValueError(f"prediction must be (B, N, 5+nc), got {tuple(pred.shape)}")
    work = pred if pred.is_contiguous() else pred.contiguous()
    batch, rows, no = work.shape
    if batch == 0:
        return ([], []) if return_rows else []
    buf = ops.get_buffers(work.device, batch, rows if cap is None else cap, no - 5)
    ops.compact_from_dense(work, conf_thres, buf, write_back=True)
    if work is not pred:
        pred[..., 4].copy_(work[..., 4])            # keep the in-place side effect on the caller's memory
    out, out_row = buf.new_outputs()
    ops.nms(buf, nms_thres, out, out_row)
    _, kept, overflow = ops.read_counts(buf)
    if overflow:
        raise ops.YoloB200Error(f"candidate capacity {buf.cap} per image exceeded; raise `cap`")
    return ops.ragged(out, out_row, kept, with_rows=return_rows)


def non_max_suppression(prediction, conf_thres=0.5, nms_thres=0.5, return_rows=False):
    """See module docstring.  ``prediction``: (B, N, 5+nc) tensor or a list of (N, 5+nc) tensors."""
    if not nms_thres < 1:
        raise ValueError("nms_thres must be < 1: the reference never terminates otherwise (utils.py:266-275)")
    if isinstance(prediction, torch.Tensor):
        return _nms_tensor(prediction, conf_thres, nms_thres, return_rows)
    dets, rows = [], []
    for pred in prediction:                           # list of per-image tensors, as the reference accepts
        if pred.dim() != 2:
            raise ValueError("list entries must be (N, 5+nc) tensors")
        r = _nms_tensor(pred.unsqueeze(0), conf_thres, nms_thres, True)
        dets.append(r[0][0])
        rows.append(r[1][0])
    return (dets, rows) if return_rows else dets
