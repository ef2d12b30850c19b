"""Load the *live* reference (Dipet/pytorch_yolo) by path -- build container only.

TEST INFRASTRUCTURE ONLY.  ``/root/reference`` exists in the build container and
not on the GPU box, so everything here is used (a) by ``tests/golden/make_golden.py``
to produce the committed golden vectors and (b) by CPU tests that skip when the
reference is absent.  Nothing is copied: the reference files are imported from
where they lie.

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

REFERENCE_ROOT = os.environ.get("YOLO_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "pytorch_yolo", "utils", "utils.py"))


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
    from pytorch_yolo.utils.utils import non_max_suppression, bbox_iou, xywh2xyxy, scale_coords, _dict_from_results
    from pytorch_yolo.models.yolov3_spp import YOLOv3SPP
    from pytorch_yolo.models.yolov3_tiny import YOLOv3Tiny
    return types.SimpleNamespace(YOLOLayer=YOLOLayer, non_max_suppression=non_max_suppression,
                                 bbox_iou=bbox_iou, xywh2xyxy=xywh2xyxy, scale_coords=scale_coords,
                                 dict_from_results=_dict_from_results,
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
