"""Generate the golden vectors under tests/golden/ from the LIVE reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

It imports the unmodified reference functions by path (oracle/ref_loader.py), runs
``YOLOLayer.forward`` + ``torch.cat`` and ``non_max_suppression`` (with the documented
stable-argsort tie rule, applied by monkeypatching ``Tensor.argsort`` -- no source
edit) on seeded inputs, and stores inputs + outputs as ``.npz``.  The vectors pin
``oracle/yolo_oracle.py`` (CPU tests) and the CUDA path (GPU tests).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_loader                      # noqa: E402
from pytorch_yolo_b200 import synth                # noqa: E402


def ref_decode(ref, heads, anchors_per_scale, nc, img_size):
    outs = []
    for h, anchors in zip(heads, anchors_per_scale):
        layer = ref.YOLOLayer(anchors, nc, anchors_per_scale).eval()
        io, _ = layer(h.clone(), img_size)
        outs.append(io)
    return torch.cat(outs, 1)


def ref_nms(ref, pred, conf, nms):
    """Returns (list of outputs, mutated prediction)."""
    pred = pred.clone()
    with ref_loader.stable_argsort():
        out = ref.non_max_suppression(pred, conf, nms)
    return out, pred


def pack_nms(out):
    counts = np.array([0 if o is None else len(o) for o in out], dtype=np.int64)
    rows = [o.numpy() for o in out if o is not None]
    flat = np.concatenate(rows, 0) if rows else np.zeros((0, 7), np.float32)
    return counts, flat.astype(np.float32)


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB")


def edge_case_prediction():
    """Hand-built (1, N, 7) nc=2 prediction exercising every branch of the filter and the MERGE loop."""
    rows = []

    def add(x, y, w, h, obj, c0, c1):
        rows.append([x, y, w, h, obj, c0, c1])

    # IoU exactly 0.5: 20x10 and 10x10 concentric (class 0)
    add(100, 100, 20, 10, 0.9, 0.9, 0.1)
    add(100, 100, 10, 10, 0.8, 0.9, 0.1)
    # width exactly 2 (dropped), just above 2 (kept), infinite width (dropped), NaN class (dropped)
    add(200, 200, 2.0, 30, 0.9, 0.9, 0.1)
    add(200, 260, 2.0001, 30, 0.9, 0.9, 0.1)
    add(300, 200, float("inf"), 30, 0.9, 0.9, 0.1)
    add(300, 260, 20, 30, 0.9, float("nan"), 0.1)
    add(float("nan"), 320, 20, 30, 0.9, 0.2, 0.9)
    # a class-1 singleton (emitted unmerged)
    add(400, 400, 40, 40, 0.7, 0.1, 0.95)
    # score exactly at the threshold 0.25 = 0.5*0.5 (strict > drops it)
    add(500, 100, 30, 30, 0.5, 0.5, 0.25)
    # class tie: first index wins
    add(500, 300, 30, 30, 0.9, 0.6, 0.6)
    # a three-box cluster + a far box (class 0): merge of 3, then last survivor unmerged
    add(50, 500, 40, 40, 0.95, 0.99, 0.0)
    add(52, 501, 40, 42, 0.85, 0.98, 0.0)
    add(49, 498, 41, 39, 0.75, 0.97, 0.0)
    add(560, 560, 25, 25, 0.31, 0.97, 0.0)
    return torch.tensor([rows], dtype=torch.float32)


def many_same_class(n=260, seed=5):
    """(1, n, 7) nc=2: all boxes class 1, > 100 of them, heavy overlaps (cap + merge)."""
    g = torch.Generator().manual_seed(seed)
    xy = 300 + 40 * torch.randn(n, 2, generator=g)
    wh = 30 + 10 * torch.rand(n, 2, generator=g)
    obj = 0.3 + 0.7 * torch.rand(n, 1, generator=g)
    c0 = 0.1 * torch.rand(n, 1, generator=g)
    c1 = 0.5 + 0.5 * torch.rand(n, 1, generator=g)
    return torch.cat((xy, wh, obj, c0, c1), 1).unsqueeze(0).float()


def main():
    torch.set_num_threads(1)
    ref = ref_loader.load()

    # --- G1: decode + NMS on small synthetic heads (odd and even plane sizes)
    for wl, kind, batch, seed, conf in (("mini-96", "B", 2, 11, 0.3), ("mini-160", "A", 1, 12, 0.05)):
        w = synth.WORKLOADS[wl]
        heads = synth.synth_heads(wl, batch, kind, seed)
        pred = ref_decode(ref, heads, w["anchors"], w["nc"], w["img_size"])
        out, mutated = ref_nms(ref, pred, conf, 0.5)
        counts, flat = pack_nms(out)
        arrays = {f"head{k}": h.numpy() for k, h in enumerate(heads)}
        save(f"decode_nms_{wl}_{kind}", workload=np.array(wl), conf=np.float64(conf), nms=np.float64(0.5),
             decoded=pred.numpy(), col4_after=mutated[..., 4].numpy(), counts=counts, dets=flat, **arrays)

    # --- G2/G3: NMS only, clustered boxes, without and with exact score ties
    for name, kw, conf, nms in (("nms_clustered", dict(batch=2, n_rows=700, nc=80, seed=21), 0.3, 0.5),
                                ("nms_ties", dict(batch=2, n_rows=500, nc=20, seed=22, tie_levels=8), 0.1, 0.45),
                                ("nms_lowconf", dict(batch=1, n_rows=900, nc=6, seed=23), 0.001, 0.5)):
        pred = synth.synth_prediction(**kw)
        out, mutated = ref_nms(ref, pred, conf, nms)
        counts, flat = pack_nms(out)
        save(name, pred=pred.numpy(), conf=np.float64(conf), nms=np.float64(nms),
             col4_after=mutated[..., 4].numpy(), counts=counts, dets=flat)

    # --- G4: more than 100 boxes in one class
    pred = many_same_class()
    out, mutated = ref_nms(ref, pred, 0.2, 0.6)
    counts, flat = pack_nms(out)
    save("nms_cap100", pred=pred.numpy(), conf=np.float64(0.2), nms=np.float64(0.6),
         col4_after=mutated[..., 4].numpy(), counts=counts, dets=flat)

    # --- G5: known-answer edge cases, at two thresholds (IoU == 0.5 exactly is NOT suppressed at 0.5)
    for nms in (0.5, 0.49):
        pred = edge_case_prediction()
        out, mutated = ref_nms(ref, pred, 0.25, nms)
        counts, flat = pack_nms(out)
        save(f"nms_edges_{int(nms * 100)}", pred=pred.numpy(), conf=np.float64(0.25), nms=np.float64(nms),
             col4_after=mutated[..., 4].numpy(), counts=counts, dets=flat)

    # --- G6: all-zero image -> None
    pred = torch.zeros(2, 50, 9)
    out, mutated = ref_nms(ref, pred, 0.1, 0.5)
    counts, flat = pack_nms(out)
    save("nms_empty", pred=pred.numpy(), conf=np.float64(0.1), nms=np.float64(0.5),
         col4_after=mutated[..., 4].numpy(), counts=counts, dets=flat)

    # --- G7: BASELINE config 1 -- YOLOv3-tiny 416x416 batch 1, random init, full reference forward
    torch.manual_seed(0)
    model = ref.YOLOv3Tiny().eval()
    x = torch.rand(1, 3, 416, 416)
    with torch.no_grad():
        b1, b2 = model._forward_encoder(x)
        pred, _ = model(x)
    out, mutated = ref_nms(ref, pred, 0.1, 0.5)
    counts, flat = pack_nms(out)
    postproc_golden(ref)
    save("tiny416_randinit", head0=b1.numpy(), head1=b2.numpy(), conf=np.float64(0.1), nms=np.float64(0.5),
         decoded_every7=pred[:, ::7].numpy(), col4_after=mutated[..., 4].numpy(), counts=counts, dets=flat)


def postproc_golden(ref):
    """G8: scale_coords + _dict_from_results (utils.py:296-327) on seeded detection rows, several letterbox geometries."""
    import json
    g = torch.Generator().manual_seed(77)
    cur_shape = (416, 608)                                   # network input (h, w), letterboxed
    orig_shapes = [(480, 640), (1080, 1920), (300, 200), (416, 608), (97, 1333)]
    dets, paths = [], []
    for i, _ in enumerate(orig_shapes):
        n = [7, 40, 1, 13, 25][i]
        xy1 = torch.rand(n, 2, generator=g) * torch.tensor([560.0, 380.0]) - 20.0     # some boxes start in the padding
        wh = 5 + 150 * torch.rand(n, 2, generator=g)
        rows = torch.cat((xy1, xy1 + wh, torch.rand(n, 2, generator=g), torch.randint(0, 80, (n, 1), generator=g).float()), 1)
        rows[::3, :4] = (rows[::3, :4] * 2).round() / 2 + 0.5                      # exact .5 values: round-half-even
        dets.append(rows.float())
        paths.append(f"img_{i}.jpg")
    dets.insert(2, None)
    paths.insert(2, "none.jpg")
    orig_shapes.insert(2, (10, 10))
    scaled = [None if d is None else ref.scale_coords(cur_shape, d[:, :4].clone(), o) for d, o in zip(dets, orig_shapes)]
    data = ref.dict_from_results({}, [None if d is None else d.clone() for d in dets], paths, orig_shapes, cur_shape)
    arrays = {}
    for i, (d, sc) in enumerate(zip(dets, scaled)):
        if d is not None:
            arrays[f"det{i}"] = d.numpy()
            arrays[f"scaled{i}"] = sc.numpy()
    save("postproc", cur_shape=np.array(cur_shape), orig_shapes=np.array(orig_shapes), n_images=np.int64(len(dets)),
         records=np.array(json.dumps(data)), **arrays)


if __name__ == "__main__":
    main()
