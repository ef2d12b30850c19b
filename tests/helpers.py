"""Shared helpers for the test-suite (golden loading, ragged unpacking, comparisons)."""
from __future__ import annotations

import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

NMS_GOLDEN = ["nms_clustered", "nms_ties", "nms_lowconf", "nms_cap100", "nms_edges_50", "nms_edges_49", "nms_empty"]
DECODE_GOLDEN = ["decode_nms_mini-96_B", "decode_nms_mini-160_A"]


def load_golden(name):
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def unpack(counts, flat):
    """counts (B,), flat (sum,7) -> list of tensors | None, like the reference's return value."""
    out, o = [], 0
    for c in counts.tolist():
        out.append(None if c == 0 else torch.from_numpy(np.ascontiguousarray(flat[o:o + c])))
        o += c
    return out


def assert_dets_equal(got, want, box_rtol=0.0, what=""):
    """Scores / class_conf / class (cols 4..6) bit-exact, order included; boxes within box_rtol (0 = bit-exact)."""
    assert len(got) == len(want), what
    for i, (g, w) in enumerate(zip(got, want)):
        if w is None:
            assert g is None, f"{what} image {i}: expected None, got {None if g is None else tuple(g.shape)}"
            continue
        assert g is not None, f"{what} image {i}: expected {tuple(w.shape)}, got None"
        g = g.detach().cpu()
        assert g.shape == w.shape, f"{what} image {i}: {tuple(g.shape)} vs {tuple(w.shape)}"
        assert g.dtype == torch.float32
        assert torch.equal(g[:, 4:], w[:, 4:]), f"{what} image {i}: score/class columns differ"
        if box_rtol == 0.0:
            assert torch.equal(g[:, :4], w[:, :4]), f"{what} image {i}: boxes differ " \
                f"(max abs {float((g[:, :4] - w[:, :4]).abs().max()):.3e})"
        else:
            torch.testing.assert_close(g[:, :4], w[:, :4], rtol=box_rtol, atol=1e-30, msg=lambda m: f"{what} image {i}: {m}")
