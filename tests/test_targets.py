"""Training-side consumers of the YOLOLayer constants against vectors the LIVE reference
produced (tests/golden/targets.npz, made by tests/golden/make_golden_targets.py).

CPU: the reference's own build_targets / compute_loss (utils/utils.py:124-197), unmodified, run on a model whose
``yolo_layers`` are THIS repo's YOLOLayer objects in training mode -- the attribute contract (n_grids, anchor_vec,
n_classes, the raw training-mode output) -- and reproduce the golden vectors bit for bit.
GPU: the one-launch ``build_targets`` kernel (csrc/targets.cu) returns the same indices and xy targets exactly and the
log-space wh targets within 1e-6; the reference's compute_loss fed by it gives the reference's loss.
"""
import types

import numpy as np
import pytest
import torch

from oracle import ref_loader
from pytorch_yolo_b200 import synth
from tests.golden.make_golden_targets import HYPER
from tests.helpers import load_golden

DEV = "cuda:0"
needs_ref = pytest.mark.skipif(not ref_loader.available(), reason="reference files not present")


def _setup(device):
    from pytorch_yolo_b200 import YOLOLayer
    g = load_golden("targets")
    wl, batch, seed = str(g["workload"]), int(g["batch"]), int(g["seed"])
    w = synth.WORKLOADS[wl]
    heads = [h.to(device) for h in synth.synth_heads(wl, batch, "A", seed)]
    layers = [YOLOLayer(a, w["nc"], w["anchors"]).train() for a in w["anchors"]]
    p = [l(h, w["img_size"]) for l, h in zip(layers, heads)]
    model = types.SimpleNamespace(yolo_layers=layers, hyper_params=dict(HYPER), n_class=w["nc"])
    return g, model, p


def _check(g, name, txy, twh, tcls, indices, wh_rtol):
    for l in range(3):
        idx = torch.stack([i.cpu() for i in indices[l]]) if len(indices[l][0]) else torch.zeros(4, 0, dtype=torch.int64)
        assert idx.dtype == torch.int64 and tcls[l].dtype == torch.int64
        assert np.array_equal(idx.numpy(), g[f"{name}_idx{l}"]), f"layer {l}: indices"
        assert np.array_equal(tcls[l].cpu().numpy(), g[f"{name}_tcls{l}"])
        assert np.array_equal(txy[l].cpu().numpy(), g[f"{name}_txy{l}"].reshape(-1, 2)), f"layer {l}: txy"
        want = torch.from_numpy(g[f"{name}_twh{l}"]).reshape(-1, 2)
        if wh_rtol == 0:
            assert torch.equal(twh[l].cpu(), want)
        else:
            torch.testing.assert_close(twh[l].cpu(), want, rtol=wh_rtol, atol=1e-6)


@needs_ref
@pytest.mark.parametrize("name", ["t", "empty"])
def test_reference_training_functions_run_on_our_layers(name):
    ref = ref_loader.load()
    g, model, p = _setup("cpu")
    targets = torch.from_numpy(g[f"{name}_targets"])
    txy, twh, tcls, indices = ref.build_targets(model, targets)
    _check(g, name, txy, twh, tcls, indices, wh_rtol=0)
    _, parts = ref.compute_loss(p, targets, model)
    assert np.array_equal(parts.numpy(), g[f"{name}_loss"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["t", "empty"])
def test_build_targets_kernel_matches_reference(name):
    from pytorch_yolo_b200.utils.targets import build_targets
    g, model, _ = _setup(DEV)
    targets = torch.from_numpy(g[f"{name}_targets"]).to(DEV)
    txy, twh, tcls, indices = build_targets(model, targets)
    _check(g, name, txy, twh, tcls, indices, wh_rtol=1e-6)
    if name == "t":
        assert sum(len(i[0]) for i in indices) > 300           # targets match anchors on more than one scale
        with pytest.raises(Exception):
            build_targets(model, targets.cpu())                 # no CPU path


@pytest.mark.gpu
@needs_ref
def test_reference_compute_loss_on_kernel_targets(monkeypatch):
    from pytorch_yolo_b200.utils.targets import build_targets
    ref = ref_loader.load()
    g, model, p = _setup(DEV)
    targets = torch.from_numpy(g["t_targets"]).to(DEV)
    monkeypatch.setattr(ref.utils_module, "build_targets", build_targets)
    _, parts = ref.compute_loss(p, targets, model)
    torch.testing.assert_close(parts.cpu(), torch.from_numpy(g["t_loss"]), rtol=1e-5, atol=1e-6)


# ---- the export-side consumer: the reference's OpenVINO RegionYolo layer (openvino_converter/layers.py:177-194) reads
# module.all_anchors / module.anchors / module.n_classes of a YOLOLayer
REGION_YOLO_DATA = {'anchors': '10,13,16,30,33,23,30,61,62,45,59,119,116,90,156,198,373,326', 'axis': '1', 'coords': '4',
                    'do_softmax': '0', 'end_axis': '3', 'mask': '0,1,2', 'num': '3', 'classes': '80'}   # from the live reference


def _region_yolo_attrs(layer):
    """What OpenVINORegionYolo.__init__ computes from the module (restated: layers.py:185-194)."""
    anchors = np.array(layer.all_anchors).flatten()
    return {'anchors': ','.join(str(int(i)) for i in anchors), 'axis': '1', 'coords': '4', 'do_softmax': '0', 'end_axis': '3',
            'mask': '0,1,2', 'num': str(len(layer.anchors)), 'classes': str(layer.n_classes)}


def test_region_yolo_export_attributes():
    from pytorch_yolo_b200 import YOLOLayer
    layer = YOLOLayer(synth.SPP_ANCHORS[1], 80, synth.SPP_ANCHORS)
    assert _region_yolo_attrs(layer) == REGION_YOLO_DATA


@pytest.mark.skipif(not __import__("os").path.isfile(ref_loader.REFERENCE_ROOT + "/pytorch_yolo/openvino_converter/layers.py"),
                    reason="reference exporter not present")
def test_reference_region_yolo_layer_accepts_our_layer():
    import importlib.util
    from pytorch_yolo_b200 import YOLOLayer
    spec = importlib.util.spec_from_file_location(
        "ref_ov_layers", ref_loader.REFERENCE_ROOT + "/pytorch_yolo/openvino_converter/layers.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    ref = ref_loader.load()
    made = [mod.OpenVINORegionYolo(7, 'yolo', 'FP32', inputs={'conv': (1, 255, 38, 38)}, module=cls(synth.SPP_ANCHORS[1], 80, synth.SPP_ANCHORS))
            for cls in (ref.YOLOLayer, YOLOLayer)]
    assert made[0].data == made[1].data == REGION_YOLO_DATA
    assert tuple(made[0].out_size) == tuple(made[1].out_size)
