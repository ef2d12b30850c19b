"""Golden vectors of the training-side consumers of the YOLOLayer constants.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_targets.py

Runs the LIVE, unmodified ``build_targets`` and ``compute_loss`` of the reference (utils/utils.py:124-197) on a stub model
whose ``yolo_layers`` are the reference's own ``YOLOLayer`` objects (three SPP-anchor scales at 160 x 160, populated by one
training-mode forward over seeded head tensors) and on a seeded target table, and stores inputs and outputs.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_loader                      # noqa: E402
from pytorch_yolo_b200 import synth                # noqa: E402

WORKLOAD, BATCH, SEED = "mini-160", 4, 91
HYPER = {'iou_thresh': 0.25, 'xy_loss': 0.5, 'wh_loss': 0.0625, 'cls_loss': 0.0625, 'conf_loss': 4.0}


def make_targets(nt: int, batch: int, nc: int, seed: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    img = torch.randint(0, batch, (nt, 1), generator=g).float()
    cls = torch.randint(0, nc, (nt, 1), generator=g).float()
    xy = torch.rand(nt, 2, generator=g) * 0.998 + 0.001
    wh = torch.exp(torch.randn(nt, 2, generator=g) * 0.9 - 2.2).clamp(0.005, 0.95)     # from a few pixels to most of the image
    t = torch.cat((img, cls, xy, wh), 1)
    t[::17, 2:4] = (torch.round(t[::17, 2:4] * 20) / 20).clamp(0.05, 0.95)                             # centres exactly on grid lines
    return t


def stub_model(layer_cls, w, heads):
    layers = [layer_cls(a, w["nc"], w["anchors"]).train() for a in w["anchors"]]
    p = [l(h, w["img_size"]) for l, h in zip(layers, heads)]                            # populates the grid constants
    return types.SimpleNamespace(yolo_layers=layers, hyper_params=dict(HYPER), n_class=w["nc"]), p


def main():
    torch.set_num_threads(1)
    ref = ref_loader.load()
    w = synth.WORKLOADS[WORKLOAD]
    heads = synth.synth_heads(WORKLOAD, BATCH, "A", SEED)
    model, p = stub_model(ref.YOLOLayer, w, heads)
    arrays = {}
    for name, nt in (("t", 300), ("empty", 0)):
        targets = make_targets(nt, BATCH, w["nc"], SEED + 1)
        txy, twh, tcls, indices = ref.build_targets(model, targets)
        arrays[f"{name}_targets"] = targets.numpy()
        for l in range(len(model.yolo_layers)):
            arrays[f"{name}_txy{l}"] = txy[l].numpy()
            arrays[f"{name}_twh{l}"] = twh[l].numpy()
            arrays[f"{name}_tcls{l}"] = np.asarray(tcls[l].numpy(), dtype=np.int64)
            arrays[f"{name}_idx{l}"] = np.stack([np.asarray(i.numpy(), dtype=np.int64) for i in indices[l]]) \
                if nt else np.zeros((4, 0), np.int64)
        loss, parts = ref.compute_loss(p, targets, model)
        arrays[f"{name}_loss"] = parts.numpy()
    path = os.path.join(HERE, "targets.npz")
    np.savez_compressed(path, workload=np.array(WORKLOAD), batch=np.int64(BATCH), seed=np.int64(SEED),
                        hyper=np.array(repr(HYPER)), **arrays)
    print(f"targets: {os.path.getsize(path) / 1024:.0f} KiB; kept per layer "
          f"{[arrays[f't_idx{l}'].shape[1] for l in range(3)]}; loss parts {arrays['t_loss']}")


if __name__ == "__main__":
    main()
