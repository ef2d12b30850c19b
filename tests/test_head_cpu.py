"""Host logic of the fused head path (no GPU): folding a reference head -- ConvBlock = Conv2d(bias=False) + BatchNorm2d +
LeakyReLU(0.1) (reference models/yolo_base.py:19-57) or a plain Conv2d (models/yolov3_tiny.py:38,42) -- into the
(weight, bias, slope) triple the kernel consumes, checked against the module itself and, when the reference is
present, against its own ConvBlock.fuse()."""
import pytest
import torch
from torch import nn

from oracle import ref_loader, yolo_oracle
from pytorch_yolo_b200 import ops
from pytorch_yolo_b200.head import find_heads, split_head


def _apply(hw, x):
    y = torch.einsum("oc,bchw->bohw", hw.weight[:hw.n_out].double(), x.double()) + hw.bias.double()[None, :, None, None]
    return torch.maximum(y, y * hw.negative_slope).float()


def test_fold_convblock_matches_module():
    torch.manual_seed(0)
    blk = nn.Sequential(nn.Conv2d(32, 255, 1, bias=False), nn.BatchNorm2d(255), nn.LeakyReLU(0.1, inplace=True))
    with torch.no_grad():
        blk[1].running_mean.normal_()
        blk[1].running_var.uniform_(0.5, 2.0)
        blk[1].weight.uniform_(0.5, 1.5)
        blk[1].bias.normal_()
    blk.eval()
    hw = ops.fold_head(blk)
    assert hw.weight.shape == (256, 32) and hw.n_out == 255 and hw.negative_slope == pytest.approx(0.1)
    assert bool((hw.weight[255] == 0).all())                       # pad row
    x = torch.randn(2, 32, 5, 7)
    with torch.no_grad():
        torch.testing.assert_close(_apply(hw, x), blk(x), rtol=1e-5, atol=1e-5)


def test_fold_plain_conv_and_rejects_other_modules():
    torch.manual_seed(1)
    conv = nn.Conv2d(64, 75, 1).eval()
    hw = ops.fold_head(conv)
    assert hw.weight.shape == (256, 64) and hw.n_out == 75 and hw.negative_slope == 1.0
    assert bool((hw.weight[75:] == 0).all())
    x = torch.randn(1, 64, 4, 4)
    with torch.no_grad():
        torch.testing.assert_close(_apply(hw, x), conv(x), rtol=1e-5, atol=1e-5)
    with pytest.raises(ValueError):
        ops.fold_head(nn.Conv2d(8, 255, 3, padding=1))             # not a 1x1 convolution
    with pytest.raises(ValueError):
        ops.fold_head(nn.Sequential(nn.Conv2d(8, 16, 1), nn.Conv2d(16, 255, 1)))


def test_split_head():
    branch = nn.Sequential(nn.Conv2d(4, 8, 3, padding=1), nn.ReLU(), nn.Conv2d(8, 18, 1))
    trunk, head = split_head(branch)
    assert len(trunk) == 2 and head is branch[2]


@pytest.mark.skipif(not ref_loader.available(), reason="reference checkout not present")
def test_fold_matches_reference_convblock_fuse():
    """The reference's own ConvBlock head (yolov3_spp.py:86) and its fuse() (yolo_base.py:46-57)."""
    ref_loader.load()
    from pytorch_yolo.models.yolo_base import ConvBlock
    torch.manual_seed(2)
    blk = ConvBlock(64, 255, size=1, stride=1)
    with torch.no_grad():
        bn = blk.sequence.batch_norm
        bn.running_mean.normal_()
        bn.running_var.uniform_(0.5, 2.0)
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.normal_()
    blk.eval()
    hw = ops.fold_head(blk)
    x = torch.randn(2, 64, 6, 6)
    with torch.no_grad():
        want = blk(x)
        torch.testing.assert_close(_apply(hw, x), want, rtol=1e-5, atol=1e-5)
        # the oracle's restatement of the head producer against the live reference module, bit for bit
        got = yolo_oracle.head_conv(x, blk.sequence.conv.weight, None, 0.1,
                                    (bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps))
        assert torch.equal(got, want)
        blk.fuse()
        fused_conv = [m for m in blk.modules() if isinstance(m, nn.Conv2d)][0]
        torch.testing.assert_close(hw.weight[:255], fused_conv.weight.view(255, 64), rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(hw.bias, fused_conv.bias, rtol=1e-5, atol=1e-5)


@pytest.mark.skipif(not ref_loader.available(), reason="reference checkout not present")
@pytest.mark.parametrize("which", ["spp", "tiny"])
def test_find_heads_on_the_reference_models(which):
    """find_heads locates the head convolutions of the live reference models; trunk + located head == _forward_encoder."""
    ref = ref_loader.load()
    torch.manual_seed(0)
    if which == "spp":
        anchors = (((116, 90), (156, 198), (373, 326)), ((30, 61), (62, 45), (59, 119)), ((10, 13), (16, 30), (33, 23)))
        model = ref.YOLOv3SPP(kernels_divider=8, anchors=anchors).eval()
        want_cin = [1024 // 8, 512 // 8, 256 // 8]
    else:
        model = ref.YOLOv3Tiny(kernels_divider=4).eval()
        want_cin = [256 // 4, 512 // 4]
    x = torch.rand(1, 3, 128, 128)
    heads = find_heads(model, x)
    assert [h[2][1] for h in heads] == want_cin
    assert all(h[3][1] == 255 for h in heads)
    with torch.no_grad():
        want = model._forward_encoder(x)
        # swap the heads out, run the trunk, apply the located head modules: identical tensors
        slots = []
        for name, mod, _, _ in heads:
            parent_name, _, attr = name.rpartition(".")
            parent = model.get_submodule(parent_name)
            slots.append((parent, attr, mod))
            setattr(parent, attr, torch.nn.Identity())
        feats = model._forward_encoder(x)
        for parent, attr, mod in slots:
            setattr(parent, attr, mod)
        for f, (_, mod, in_shape, _), w in zip(feats, heads, want):
            assert tuple(f.shape) == in_shape
            assert torch.equal(mod(f), w)


def test_ctypes_structs_match_the_c_header(tmp_path):
    """The ctypes mirrors in _lib.py against include/yolo_b200.h as a C compiler lays it out (size and every offset)."""
    import ctypes
    import os
    import shutil
    import subprocess
    from pytorch_yolo_b200 import _lib
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    fields = {"yolo_b200_scale": ["head", "ny", "nx", "na", "row_off", "stride", "anchor_vec"],
              "yolo_b200_head": ["x", "weight", "bias_host", "head_out", "c_in", "x_row_pitch", "negative_slope", "scale"]}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "yolo_b200.h"', 'int main(void) {']
    for name, fs in fields.items():
        lines.append(f'  printf("{name} %zu", sizeof({name}));')
        for f in fs:
            lines.append(f'  printf(" %zu", offsetof({name}, {f}));')
        lines.append('  printf("\\n");')
    lines += ['  return 0;', '}']
    src = tmp_path / "abi.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "abi"
    subprocess.run([cc, "-I", os.path.join(root, "include"), "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n")
    mirrors = {"yolo_b200_scale": _lib.Scale, "yolo_b200_head": _lib.Head}
    for line in filter(None, out):
        name, size, *offs = line.split()
        st = mirrors[name]
        assert ctypes.sizeof(st) == int(size), name
        assert [getattr(st, f).offset for f in fields[name]] == [int(o) for o in offs], name


def test_fold_head_after_convblock_fuse_and_with_non_affine_bn():
    """ConvBlock.fuse() (reference yolo_base.py:46-57) leaves Conv2d(bias) + LeakyReLU; a BatchNorm without affine
    parameters is legal torch too: both fold to the same function as the module."""
    torch.manual_seed(5)
    fused_like = nn.Sequential(nn.Conv2d(32, 255, 1, bias=True), nn.LeakyReLU(0.1, inplace=True)).eval()
    no_affine = nn.Sequential(nn.Conv2d(32, 255, 1, bias=False), nn.BatchNorm2d(255, affine=False), nn.LeakyReLU(0.1)).eval()
    with torch.no_grad():
        no_affine[1].running_mean.normal_()
        no_affine[1].running_var.uniform_(0.5, 2.0)
    x = torch.randn(2, 32, 3, 5)
    for mod in (fused_like, no_affine):
        hw = ops.fold_head(mod)
        assert hw.negative_slope == pytest.approx(0.1) and hw.weight.shape == (256, 32)
        with torch.no_grad():
            torch.testing.assert_close(_apply(hw, x), mod(x), rtol=1e-5, atol=1e-5)


def test_fold_head_property_random_modules():
    """Randomised: channel counts, slopes, BatchNorm on / off, bias on / off -- the folded triple reproduces the module."""
    g = torch.Generator().manual_seed(11)
    for trial in range(12):
        c_in = int(torch.randint(1, 9, (1,), generator=g)) * 8
        n_out = int(torch.randint(1, 256, (1,), generator=g))
        with_bn = bool(torch.randint(0, 2, (1,), generator=g))
        slope = float(torch.rand(1, generator=g))
        layers = [nn.Conv2d(c_in, n_out, 1, bias=not with_bn or trial % 3 == 0)]
        if with_bn:
            layers.append(nn.BatchNorm2d(n_out))
        if trial % 4:
            layers.append(nn.LeakyReLU(slope))
        mod = nn.Sequential(*layers).eval()
        with torch.no_grad():
            for p_ in mod.parameters():
                p_.copy_(torch.randn(p_.shape, generator=g))
            if with_bn:
                mod[1].running_mean.copy_(torch.randn(n_out, generator=g))
                mod[1].running_var.copy_(torch.rand(n_out, generator=g) + 0.25)
                mod[1].weight.copy_(torch.rand(n_out, generator=g) + 0.5)
        hw = ops.fold_head(mod)
        assert hw.n_out == n_out and hw.c_in == c_in and bool((hw.weight[n_out:] == 0).all())
        assert hw.negative_slope == (pytest.approx(slope) if trial % 4 else 1.0)
        x = torch.randn(1, c_in, 2, 3, generator=g)
        with torch.no_grad():
            torch.testing.assert_close(_apply(hw, x), mod(x), rtol=2e-5, atol=2e-5)
