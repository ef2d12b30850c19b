"""The fused hot path: heads -> detections, without the dense (B, N, 5+nc) tensor.

``detect`` is the one-shot call; ``Detector`` keeps the device buffers (and optionally a captured
CUDA graph) for repeated batches of one shape, which is how a serving / evaluation loop such as the
reference's ``test_model`` (utils/utils.py:357-393) would drive it.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from . import ops


def detect(heads: Sequence[torch.Tensor], specs: Sequence[ops.ScaleSpec], nc: int,
           conf_thres: float = 0.5, nms_thres: float = 0.5, cap: Optional[int] = None,
           return_rows: bool = False):
    """decode + filter + compaction + NMS; returns the reference's list of (n,7) tensors / None."""
    if not nms_thres < 1:
        raise ValueError("nms_thres must be < 1: the reference never terminates otherwise (utils.py:266-275)")
    batch = heads[0].shape[0]
    if batch == 0:
        return ([], []) if return_rows else []
    rows = sum(s.rows for s in specs)
    buf = ops.get_buffers(heads[0].device, batch, rows if cap is None else min(cap, rows), nc)
    ops.decode_compact(heads, specs, nc, conf_thres, buf)
    out, out_row = buf.new_outputs()
    ops.nms(buf, nms_thres, out, out_row)
    _, kept, overflow = ops.read_counts(buf)
    if overflow:
        raise ops.YoloB200Error(f"candidate capacity {buf.cap} per image exceeded; raise `cap`")
    return ops.ragged(out, out_row, kept, with_rows=return_rows)
