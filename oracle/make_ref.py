"""Recipe for ``oracle/_ref/``: the UNMODIFIED reference files of the hot path, so the real reference can be timed
on the GPU box's host cores (``bench.py --impl reference`` and the ``cpu_baseline`` leg, ``kind: "reference"``).

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  ``/root/reference`` exists in the build container and not on the GPU box;
``oracle/_ref/`` is git-ignored (no reference source enters the history) but travels with the gpurun snapshot, exactly
like the built ``.so``.  The reference's own packaging cannot be used: ``setup.py`` reads a non-existent file
 and ``import pytorch_yolo`` fails, so the files are taken where they lie and
loaded by path (``oracle/ref_loader.py``).  Files are copied byte for byte; a manifest with their
SHA-256 is written next to them and checked at load time.

    python -m oracle.make_ref            # (re)creates oracle/_ref/ from /root/reference

Called by ``__graft_entry__.build()`` whenever the reference checkout is present.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SOURCE = os.environ.get("YOLO_REFERENCE_ROOT", "/root/reference")

# the hot path (models/yolo_layer.py, utils/utils.py), what utils.py imports at module level, and the two models whose
# forward produces the head tensors of the BASELINE configs (cfg 1 runs the tiny model's full forward on the CPU)
FILES = [
    "pytorch_yolo/models/__init__.py",
    "pytorch_yolo/models/yolo_layer.py",
    "pytorch_yolo/models/yolo_base.py",
    "pytorch_yolo/models/yolov3_tiny.py",
    "pytorch_yolo/models/yolov3_spp.py",
    "pytorch_yolo/utils/__init__.py",
    "pytorch_yolo/utils/utils.py",
    "pytorch_yolo/utils/coco_helper.py",
    "pytorch_yolo/utils/torch_utils.py",
    "pytorch_yolo/openvino_converter/layers.py",     # the export-side consumer of YOLOLayer attributes (tests/test_targets.py)
]


def _sha(path: str) -> str:
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def source_available() -> bool:
    return all(os.path.isfile(os.path.join(SOURCE, f)) for f in FILES)


def make(verbose: bool = False) -> str:
    """Copy FILES from the reference checkout into oracle/_ref/ (idempotent) and write the manifest."""
    if not source_available():
        raise FileNotFoundError(f"reference checkout not found under {SOURCE}")
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(SOURCE, rel), os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not os.path.isfile(dst) or _sha(dst) != _sha(src):
            shutil.copyfile(src, dst)
        manifest[rel] = _sha(dst)
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as f:
        json.dump({"source": "Dipet/pytorch_yolo (unmodified files)", "sha256": manifest}, f, indent=1)
    if verbose:
        print(f"oracle/_ref: {len(manifest)} reference files")
    return DEST


def verify() -> bool:
    """True when oracle/_ref/ holds every file of the manifest, unmodified."""
    try:
        with open(os.path.join(DEST, "MANIFEST.json")) as f:
            want = json.load(f)["sha256"]
    except Exception:  # noqa: BLE001
        return False
    return set(want) == set(FILES) and all(
        os.path.isfile(os.path.join(DEST, r)) and _sha(os.path.join(DEST, r)) == h for r, h in want.items())


if __name__ == "__main__":
    make(verbose=True)
    sys.exit(0 if verify() else 1)
