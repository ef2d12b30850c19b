"""Load the *live* reference (Dipet/pytorch_yolo) by path.

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  ``/root/reference`` exists in the build container and
not on the GPU box.  Used (a) by ``tests/golden/make_golden.py`` to produce the committed golden
vectors, (b) by CPU tests that skip when the reference is absent and (c) by ``bench.py``'s
``--impl reference`` arm / ``cpu_baseline`` leg, which time the unmodified reference functions on
the host cores.  The files are imported from where they lie: ``/root/reference`` in the build
container, else the byte-for-byte copy ``oracle/_ref/`` that ``oracle/make_ref.py`` writes
(git-ignored, shipped to the GPU box by gpurun).

``import pytorch_yolo`` itself fails (torchvision API drift,
missing efficientnet_pytorch / tensorboardX / pycocotools, broken intra-package
imports), so the package object is pre-seeded and only the hot-path modules are
imported.
"""
from __future__ import annotations

import contextlib
import os
import sys
import types

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))


def _has(root: str) -> bool:
    return bool(root) and os.path.isfile(os.path.join(root, "pytorch_yolo", "utils", "utils.py"))


def reference_root() -> str:
    """First existing of: $YOLO_REFERENCE_ROOT, /root/reference, oracle/_ref (the shipped copy).  '' when none."""
    for cand in (os.environ.get("YOLO_REFERENCE_ROOT"), "/root/reference", os.path.join(_HERE, "_ref")):
        if _has(cand):
            return cand
    return ""


REFERENCE_ROOT = reference_root()


def available() -> bool:
    return _has(REFERENCE_ROOT)


def _seed_modules():
    if "pytorch_yolo" not in sys.modules or not hasattr(sys.modules["pytorch_yolo"], "__path__"):
        pkg = types.ModuleType("pytorch_yolo")
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "pytorch_yolo")]
        sys.modules["pytorch_yolo"] = pkg
    for name, attrs in (("pycocotools", ()), ("pycocotools.coco", ("COCO",)),
                        ("pycocotools.cocoeval", ("COCOeval",))):
        if name not in sys.modules:
            m = types.ModuleType(name)
            for a in attrs:
                setattr(m, a, type(a, (), {}))
            sys.modules[name] = m


def load():
    """Returns a namespace with the reference's YOLOLayer, non_max_suppression, bbox_iou,
    xywh2xyxy, scale_coords, YOLOv3SPP, YOLOv3Tiny."""
    if not available():
        raise FileNotFoundError(f"reference not found under {REFERENCE_ROOT}")
    _seed_modules()
    from pytorch_yolo.models.yolo_layer import YOLOLayer
    from pytorch_yolo.utils.utils import (non_max_suppression, bbox_iou, xywh2xyxy, scale_coords, _dict_from_results,
                                          build_targets, compute_loss)
    import pytorch_yolo.utils.utils as utils_module
    from pytorch_yolo.models.yolov3_spp import YOLOv3SPP
    from pytorch_yolo.models.yolov3_tiny import YOLOv3Tiny
    return types.SimpleNamespace(YOLOLayer=YOLOLayer, non_max_suppression=non_max_suppression,
                                 bbox_iou=bbox_iou, xywh2xyxy=xywh2xyxy, scale_coords=scale_coords,
                                 dict_from_results=_dict_from_results, build_targets=build_targets,
                                 compute_loss=compute_loss, utils_module=utils_module,
                                 YOLOv3SPP=YOLOv3SPP, YOLOv3Tiny=YOLOv3Tiny)


@contextlib.contextmanager
def stable_argsort():
    """Force ``Tensor.argsort`` to be stable while the reference runs (no source edit).

    This is the documented tie rule: equal scores keep ascending anchor-row order.
    """
    orig = torch.Tensor.argsort

    def patched(self, dim=-1, descending=False, stable=False):
        return orig(self, dim=dim, descending=descending, stable=True)

    torch.Tensor.argsort = patched
    try:
        yield
    finally:
        torch.Tensor.argsort = orig
