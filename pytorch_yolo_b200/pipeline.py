"""Evaluation / serving loop around the hot path: the part of the reference's
``test_model`` (utils/utils.py:357-393) between the data loader and the COCO scorer, with its stages overlapped.

Reference loop, per batch and strictly in sequence: ``imgs.to(device)`` (:374) -> ``model(imgs)`` (:376) ->
``non_max_suppression`` (:378) -> ``_dict_from_results`` (:379, a per-row Python loop behind a D2H copy).

Here, with ``depth`` batches in flight:

* the next batch's images go host -> device on a copy stream while the current batch computes;
* the backbone (the reference's own ``_forward_encoder`` -- out of scope, run as is) feeds the fused
  decode + NMS kernels directly, no dense (B, N, 5+nc) tensor, no ``torch.cat``;
* un-letterboxing + rounding run as one kernel over the whole batch (``Detector.scale_to_original``);
* the kept rows come back in one pinned D2H copy per batch and are turned into records while the GPU is already
  working on the following batches.

The data set, its collate function and the COCO scoring stay the reference's.
"""
from __future__ import annotations

from collections import deque
from typing import Dict, Iterable, List, Sequence, Tuple

import torch

from . import ops
from .detect import Detector


class EvalPipeline:
    """``model`` must expose ``_forward_encoder(x) -> tuple of head tensors`` and ``yolo_layers`` (the reference's
    models do: models/yolov3_spp.py:119-139, :166-168), with :class:`pytorch_yolo_b200.YOLOLayer` heads."""

    def __init__(self, model, device, conf_thresh: float = 0.1, nms_thresh: float = 0.1, depth: int = 2):
        self.model = model
        self.device = torch.device(device)
        self.conf, self.nms = conf_thresh, nms_thresh        # test_model defaults (utils.py:359)
        self.depth = max(1, depth)
        self.copy_stream = torch.cuda.Stream(self.device)
        self.streams = [torch.cuda.Stream(self.device) for _ in range(self.depth)]
        self._detectors: Dict[tuple, List[Detector]] = {}

    def _detector(self, lane: int, heads: Sequence[torch.Tensor], img_size) -> Detector:
        layers = self.model.yolo_layers
        key = (tuple(tuple(h.shape) for h in heads), img_size)
        lanes = self._detectors.get(key)
        if lanes is None:
            specs = [l._prepare(h, img_size) for l, h in zip(layers, heads)]
            nc = layers[0].n_classes
            # eager launches: the backbone hands over freshly allocated head tensors every batch
            lanes = [Detector(specs, nc, heads[0].shape[0], self.device, self.conf, self.nms, use_graph=False)
                     for _ in range(self.depth)]
            self._detectors[key] = lanes
        return lanes[lane]

    @torch.no_grad()
    def run(self, batches: Iterable[Tuple[torch.Tensor, Sequence[str], Sequence[Tuple[int, int]]]],
            data: dict | None = None) -> dict:
        """``batches`` yields ``(imgs (B,3,H,W) CPU tensor, image paths, original (h, w) shapes)``; returns the
        ``{path: [{'type','score','left','top','right','bottom'}, ...]}`` dict ``_dict_from_results`` builds."""
        data = {} if data is None else data
        pending = deque()
        it = iter(batches)
        step = 0

        def upload(batch):
            imgs, paths, shapes = batch
            src = imgs if imgs.is_pinned() else imgs.pin_memory()
            with torch.cuda.stream(self.copy_stream):
                dev = src.to(self.device, non_blocking=True)            # utils.py:374
                ev = torch.cuda.Event()
                ev.record(self.copy_stream)
            return dev, src, ev, paths, shapes

        nxt = next(it, None)
        staged = upload(nxt) if nxt is not None else None
        while staged is not None:
            dev_imgs, keep_src, ev, paths, shapes = staged
            nxt = next(it, None)
            staged = upload(nxt) if nxt is not None else None             # overlaps with the compute below
            lane = step % self.depth
            st = self.streams[lane]
            with torch.cuda.stream(st):
                st.wait_event(ev)
                dev_imgs.record_stream(st)
                heads = self.model._forward_encoder(dev_imgs)              # utils.py:376 (backbone, as is)
                cur_shape = tuple(dev_imgs.shape[-2:])
                det = self._detector(lane, heads, max(cur_shape))
                det.launch(list(heads))
            pending.append((det, st, paths, shapes, cur_shape, heads))
            step += 1
            if len(pending) >= self.depth:
                self._collect(pending.popleft(), data)
        while pending:
            self._collect(pending.popleft(), data)
        return data

    def _collect(self, item, data: dict) -> None:
        det, st, paths, shapes, cur_shape, _heads = item
        with torch.cuda.stream(st):
            _, kept = det.counts()                                          # waits for this batch only
            n_max = int(kept.max()) if len(kept) else 0
            if n_max == 0:
                return
            det.scale_to_original(cur_shape, shapes, do_round=True)        # utils.py:313
            rows = det.out[:, :n_max].to("cpu", non_blocking=False)
        rows = rows.tolist()
        for i, n in enumerate(kept.tolist()):
            if n:
                recs = [{'type': int(r[6]), 'score': float(r[4]), 'left': int(r[0]), 'top': int(r[1]),
                         'right': int(r[2]), 'bottom': int(r[3])} for r in rows[i][:n]]
                data.setdefault(paths[i], []).extend(recs)                  # utils.py:314-325
