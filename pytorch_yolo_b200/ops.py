"""Torch-facing wrappers over the C ABI (``include/yolo_b200.h``).

PyTorch is plumbing here: it owns device memory and streams; every computation is a kernel in
``libyolo_b200.so``.  There is no CPU path -- tensors that are not on a CUDA device raise.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import DET_COLS, MAX_ANCHORS, MAX_SCALES, Scale, YoloB200Error, check

MIN_WH = 2.0           # reference utils/utils.py:207
MAX_PER_CLASS = 100    # reference utils/utils.py:247-250


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{what}: expected a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise YoloB200Error(f"{what} is on {t.device}: pytorch_yolo_b200 has no CPU path "
                            "(the kernels are sm_100a CUDA); move the tensor to a CUDA device")
    if t.dtype != torch.float32:
        raise TypeError(f"{what}: expected float32, got {t.dtype}")


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


@dataclass
class ScaleSpec:
    """Host-side description of one YOLO layer (reference models/yolo_layer.py:101-111)."""
    anchors: Tuple[Tuple[float, float], ...]
    ny: int
    nx: int
    stride: float
    anchor_vec: torch.Tensor      # (na, 2) fp32 on CPU = anchors / stride, computed like the reference does

    @property
    def na(self) -> int:
        return len(self.anchors)

    @property
    def rows(self) -> int:
        return self.na * self.ny * self.nx


def scale_spec(anchors, ny: int, nx: int, img_size) -> ScaleSpec:
    stride = img_size / max(nx, ny)                                            # yolo_layer.py:102
    vec = torch.tensor(anchors, dtype=torch.float32).view(-1, 2) / stride       # yolo_layer.py:109
    if vec.shape[0] > MAX_ANCHORS:
        raise ValueError(f"at most {MAX_ANCHORS} anchors per scale are supported, got {vec.shape[0]}")
    return ScaleSpec(tuple((float(a), float(b)) for a, b in anchors), int(ny), int(nx), float(stride), vec)


def _fill_scales(heads: Sequence[torch.Tensor], specs: Sequence[ScaleSpec], nc: int):
    if not 1 <= len(heads) <= MAX_SCALES or len(heads) != len(specs):
        raise ValueError(f"need 1..{MAX_SCALES} heads with one spec each, got {len(heads)} / {len(specs)}")
    arr = (Scale * len(heads))()
    batch = heads[0].shape[0]
    dev = heads[0].device
    row_off = 0
    keep = []
    for k, (h, sp) in enumerate(zip(heads, specs)):
        _require_cuda(h, f"head {k}")
        if h.device != dev or h.shape[0] != batch:
            raise ValueError("all heads must share device and batch size")
        if h.dim() != 4 or h.shape[1] != sp.na * (nc + 5) or h.shape[2] != sp.ny or h.shape[3] != sp.nx:
            raise ValueError(f"head {k}: expected (B, {sp.na * (nc + 5)}, {sp.ny}, {sp.nx}), got {tuple(h.shape)}")
        if not h.is_contiguous():
            h = h.contiguous()          # the reference's .view at yolo_layer.py:67 needs contiguity as well
        keep.append(h)
        s = arr[k]
        s.head = h.data_ptr()
        s.ny, s.nx, s.na, s.row_off, s.stride = sp.ny, sp.nx, sp.na, row_off, sp.stride
        av = sp.anchor_vec.tolist()
        for a in range(sp.na):
            s.anchor_vec[a][0], s.anchor_vec[a][1] = av[a][0], av[a][1]
        row_off += sp.rows
    return arr, keep, batch, row_off, dev


# ----------------------------------------------------------------------------------------------
def decode_dense(heads: Sequence[torch.Tensor], specs: Sequence[ScaleSpec], nc: int,
                 out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """All scales -> one (B, N, 5+nc) tensor (YOLOLayer.forward eval + torch.cat), one launch."""
    lib = _lib.load()
    arr, keep, batch, rows, dev = _fill_scales(heads, specs, nc)
    if out is None:
        out = torch.empty(batch, rows, nc + 5, dtype=torch.float32, device=dev)
    elif out.shape != (batch, rows, nc + 5) or not out.is_contiguous() or out.device != dev:
        raise ValueError("out must be a contiguous (B, N, 5+nc) tensor on the heads' device")
    with torch.cuda.device(dev):
        check(lib.yolo_b200_decode_dense(arr, len(keep), batch, nc, rows, out.data_ptr(), _stream_ptr(dev)),
              "yolo_b200_decode_dense")
    return out


class Buffers:
    """Device buffers of one (batch, capacity, classes) problem: candidates, counters, NMS workspace.

    Everything the C ABI needs is caller-owned; this object is that caller.  ``meta`` holds, in one
    int32 tensor, ``count[B] | overflow[1] | out_count[B]`` so a single D2H copy reads all of it (and one
    memset zeroes count + overflow).
    """

    def __init__(self, device, batch: int, cap: int, nc: int, max_per_class: int = MAX_PER_CLASS):
        lib = _lib.load()
        self.device = torch.device(device)
        self.batch, self.cap, self.nc, self.mpc = batch, cap, nc, max_per_class
        self.out_cap = min(cap, nc * max_per_class)
        self.cand_box = torch.empty(max(1, batch * cap), 4, dtype=torch.float32, device=self.device)
        self.cand_meta = torch.empty(max(1, batch * cap), 4, dtype=torch.int32, device=self.device)
        self.meta = torch.zeros(2 * batch + 1, dtype=torch.int32, device=self.device)
        ws = lib.yolo_b200_nms_workspace_bytes(batch, cap, nc, max_per_class)
        self.workspace = torch.empty(max(256, ws), dtype=torch.uint8, device=self.device)
        self.meta_host = torch.empty(2 * batch + 1, dtype=torch.int32).pin_memory()

    @property
    def count_ptr(self) -> int:
        return self.meta.data_ptr()

    @property
    def overflow_ptr(self) -> int:
        return self.meta.data_ptr() + 4 * self.batch

    @property
    def out_count_ptr(self) -> int:
        return self.meta.data_ptr() + 4 * (self.batch + 1)

    def new_outputs(self):
        out = torch.empty(self.batch, self.out_cap, DET_COLS, dtype=torch.float32, device=self.device)
        out_row = torch.empty(self.batch, self.out_cap, dtype=torch.int32, device=self.device)
        return out, out_row


_buffer_cache = {}


def get_buffers(device, batch: int, cap: int, nc: int, max_per_class: int = MAX_PER_CLASS) -> Buffers:
    key = (torch.device(device), batch, cap, nc, max_per_class)
    buf = _buffer_cache.get(key)
    if buf is None:
        if len(_buffer_cache) > 8:
            _buffer_cache.clear()
        buf = _buffer_cache[key] = Buffers(device, batch, cap, nc, max_per_class)
    return buf


DECODE_VARIANTS = {"auto": 0, "ldg": 1, "tma": 2, "tma2d": 3}


def decode_compact(heads, specs, nc: int, conf_thres: float, buf: Buffers, min_wh: float = MIN_WH,
                   variant: str = "auto") -> None:
    """Fused decode + filter + compaction into ``buf`` (no (B, N, 5+nc) tensor is materialised).
    ``variant``: "auto" | "ldg" | "tma" | "tma2d" (identical results; see include/yolo_b200.h)."""
    lib = _lib.load()
    arr, keep, batch, rows, dev = _fill_scales(heads, specs, nc)
    if batch != buf.batch or nc != buf.nc or dev != buf.device:
        raise ValueError("buffer does not match the problem")
    with torch.cuda.device(dev):
        check(lib.yolo_b200_decode_compact_ex(arr, len(keep), batch, nc, rows, conf_thres, min_wh,
                                              buf.cand_box.data_ptr(), buf.cand_meta.data_ptr(), buf.cap,
                                              buf.count_ptr, buf.overflow_ptr, DECODE_VARIANTS[variant], _stream_ptr(dev)),
              "yolo_b200_decode_compact")


def compact_from_dense(pred: torch.Tensor, conf_thres: float, buf: Buffers, write_back: bool = True,
                       min_wh: float = MIN_WH) -> None:
    """Filter + compaction of a decoded (B, N, 5+nc) tensor; writes obj*cls back into column 4 like the reference."""
    lib = _lib.load()
    _require_cuda(pred, "prediction")
    if pred.dim() != 3 or not pred.is_contiguous():
        raise ValueError("prediction must be a contiguous (B, N, 5+nc) tensor")
    batch, rows, no = pred.shape
    if batch != buf.batch or no - 5 != buf.nc or pred.device != buf.device:
        raise ValueError("buffer does not match the problem")
    with torch.cuda.device(pred.device):
        check(lib.yolo_b200_compact_from_dense(pred.data_ptr(), batch, rows, no - 5, conf_thres, min_wh,
                                               1 if write_back else 0,
                                               buf.cand_box.data_ptr(), buf.cand_meta.data_ptr(), buf.cap,
                                               buf.count_ptr, buf.overflow_ptr, _stream_ptr(pred.device)),
              "yolo_b200_compact_from_dense")


def nms(buf: Buffers, nms_thres: float, out: torch.Tensor, out_row: torch.Tensor,
        out_ptrs: Optional[Tuple[int, int, int]] = None) -> None:
    """Segmented MERGE-NMS of the candidates in ``buf``.  ``out_ptrs`` overrides the destination
    (out, out_row, out_count) with raw device pointers, e.g. a peer GPU's buffers."""
    lib = _lib.load()
    if not nms_thres < 1.0:
        raise ValueError("nms_thres must be < 1: the reference never terminates otherwise (utils.py:266-275)")
    if out_ptrs is None:
        if out.shape[0] != buf.batch or out.shape[1] < buf.out_cap or not out.is_contiguous():
            raise ValueError("out must be (B, >= out_cap, 7) contiguous")
        out_ptrs = (out.data_ptr(), out_row.data_ptr(), buf.out_count_ptr)
        out_cap = out.shape[1]
    else:
        out_cap = buf.out_cap
    with torch.cuda.device(buf.device):
        check(lib.yolo_b200_nms(buf.cand_box.data_ptr(), buf.cand_meta.data_ptr(), buf.count_ptr,
                                buf.batch, buf.cap, buf.nc, nms_thres, buf.mpc,
                                out_ptrs[0], out_ptrs[1], out_cap, out_ptrs[2],
                                buf.workspace.data_ptr(), buf.workspace.numel(), _stream_ptr(buf.device)),
              "yolo_b200_nms")


def read_counts(buf: Buffers):
    """The one D2H of the path: candidate counts, kept counts and the overflow flag."""
    buf.meta_host.copy_(buf.meta, non_blocking=True)
    torch.cuda.current_stream(buf.device).synchronize()
    m = buf.meta_host
    b = buf.batch
    return m[:b], m[b + 1:2 * b + 1], int(m[b])


def ragged(out: torch.Tensor, out_row: Optional[torch.Tensor], kept_counts, with_rows: bool = False):
    """(B, out_cap, 7) + counts -> the reference's return value: list of (n, 7) tensors or None."""
    dets: List[Optional[torch.Tensor]] = []
    rows: List[Optional[torch.Tensor]] = []
    for i, n in enumerate(kept_counts.tolist()):
        dets.append(out[i, :n] if n else None)
        if with_rows:
            rows.append(out_row[i, :n] if n else None)
    return (dets, rows) if with_rows else dets
