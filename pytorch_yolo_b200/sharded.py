"""Multi-GPU: image batches shard across ranks, kept rows gather on the root over NVLink peer stores.

The reference has no multi-GPU path.  Images are independent, so the path shards
with no data-path collective: one process per GPU (``torch.distributed`` is used only to exchange a
64-byte CUDA-IPC handle at set-up and for barriers), every rank runs decode_compact + NMS on its own
contiguous slice of the batch, and its ``nms_finalize_kernel`` writes the kept rows of image ``g``
straight into ``root_out[g]`` / ``root_count[g]`` -- the root GPU's memory, mapped through CUDA IPC --
so the "ragged gather" is a handful of coalesced peer stores per image instead of a collective.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from . import _lib, ops

HANDLE_BYTES = 64


@dataclass(frozen=True)
class ShardPlan:
    """Contiguous image slices [start, end) per rank; the first ``global_batch % world`` ranks get one extra."""
    global_batch: int
    world: int

    def bounds(self, rank: int) -> Tuple[int, int]:
        base, extra = divmod(self.global_batch, self.world)
        start = rank * base + min(rank, extra)
        return start, start + base + (1 if rank < extra else 0)

    def local_batch(self, rank: int) -> int:
        s, e = self.bounds(rank)
        return e - s

    def owner(self, image: int) -> int:
        for r in range(self.world):
            s, e = self.bounds(r)
            if s <= image < e:
                return r
        raise IndexError(image)


@dataclass(frozen=True)
class GatherLayout:
    """Byte layout of the root's result buffer: out (B, out_cap, 7) f32 | out_row (B, out_cap) i32 | out_count (B) i32."""
    global_batch: int
    out_cap: int

    @property
    def out_off(self) -> int:
        return 0

    @property
    def row_off(self) -> int:
        return _align(self.global_batch * self.out_cap * ops.DET_COLS * 4)

    @property
    def count_off(self) -> int:
        return self.row_off + _align(self.global_batch * self.out_cap * 4)

    @property
    def total(self) -> int:
        return self.count_off + _align(self.global_batch * 4)

    def slice_ptrs(self, base: int, first_image: int) -> Tuple[int, int, int]:
        """Device pointers (out, out_row, out_count) of the slice starting at global image ``first_image``."""
        return (base + self.out_off + first_image * self.out_cap * ops.DET_COLS * 4,
                base + self.row_off + first_image * self.out_cap * 4,
                base + self.count_off + first_image * 4)


def _align(n: int, a: int = 256) -> int:
    return (n + a - 1) // a * a


def exchange_handle(handle: Optional[bytes], root: int = 0, group=None) -> bytes:
    """Broadcast the root's 64-byte IPC handle to every rank (works on any backend; gloo in the CPU tests)."""
    box = [handle if dist.get_rank(group) == root else None]
    dist.broadcast_object_list(box, src=dist.get_global_rank(group, root) if group is not None else root, group=group)
    got = box[0]
    if not isinstance(got, (bytes, bytearray)) or len(got) != HANDLE_BYTES:
        raise ops.YoloB200Error("peer handle exchange failed")
    return bytes(got)


class _DevMem:
    """Raw device allocation exposed to torch through __cuda_array_interface__ (no copy)."""

    def __init__(self, ptr: int, nbytes: int):
        self.ptr, self.nbytes = ptr, nbytes
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class RootGather:
    """The result buffer on the root rank (``depth`` copies, one per pipeline lane) and its peer mapping on
    every other rank."""

    def __init__(self, layout: GatherLayout, device, root: int = 0, group=None, depth: int = 1):
        lib = _lib.load()
        self.layout, self.root, self.group, self.depth = layout, root, group, depth
        self.rank = dist.get_rank(group)
        self.device = torch.device(device)
        self._owned = self._mapped = None
        total = layout.total * depth
        with torch.cuda.device(self.device):
            if self.rank == root:
                p = C.c_void_p()
                _lib.check(lib.yolo_b200_device_alloc(total, C.byref(p)), "yolo_b200_device_alloc")
                self._owned = p.value
                h = (C.c_ubyte * HANDLE_BYTES)()
                _lib.check(lib.yolo_b200_peer_export(p, h), "yolo_b200_peer_export")
                exchange_handle(bytes(h), root, group)
                self.base = p.value
                self._mem = _DevMem(self.base, total)
                self.bytes = torch.as_tensor(self._mem, device=self.device)
                self.bytes.zero_()
                torch.cuda.synchronize(self.device)
            else:
                handle = exchange_handle(None, root, group)
                h = (C.c_ubyte * HANDLE_BYTES).from_buffer_copy(handle)
                p = C.c_void_p()
                _lib.check(lib.yolo_b200_peer_open(h, C.byref(p)), "yolo_b200_peer_open")
                self._mapped = p.value
                self.base = p.value
        dist.barrier(group=group)

    def lane_base(self, lane: int) -> int:
        return self.base + lane * self.layout.total

    def root_views(self, lane: int = 0):
        """(out, out_row, out_count) tensors over copy ``lane`` of the root buffer -- root rank only."""
        L = self.layout
        b = self.bytes[lane * L.total:(lane + 1) * L.total]
        out = b[L.out_off:L.out_off + L.global_batch * L.out_cap * ops.DET_COLS * 4].view(torch.float32)
        row = b[L.row_off:L.row_off + L.global_batch * L.out_cap * 4].view(torch.int32)
        cnt = b[L.count_off:L.count_off + L.global_batch * 4].view(torch.int32)
        return out.view(L.global_batch, L.out_cap, ops.DET_COLS), row.view(L.global_batch, L.out_cap), cnt

    def close(self):
        lib = _lib.load()
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)
            dist.barrier(group=self.group)
            if self._mapped:
                lib.yolo_b200_peer_close(C.c_void_p(self._mapped)); self._mapped = None
            dist.barrier(group=self.group)
            if self._owned:
                self.bytes = None
                lib.yolo_b200_device_free(C.c_void_p(self._owned)); self._owned = None


class ShardedDetector:
    """One per rank.  ``submit(local_heads)`` runs the fused path on this rank's image slice (pipelined over
    ``depth`` streams) and stores its kept rows into the root's buffer; ``wait(ticket)`` waits for the local part
    of that step; ``gather(ticket)`` (collective: barrier) returns the global ragged list on the root and ``None``
    elsewhere."""

    def __init__(self, specs: Sequence[ops.ScaleSpec], nc: int, global_batch: int, device,
                 conf_thres: float = 0.5, nms_thres: float = 0.5, root: int = 0, group=None,
                 use_graph: bool = True, depth: int = 2, variant: str = "auto"):
        from .detect import PipelinedDetector
        self.group, self.root, self.depth = group, root, depth
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.plan = ShardPlan(global_batch, self.world)
        self.first, self.last = self.plan.bounds(self.rank)
        rows = sum(s.rows for s in specs)
        out_cap = min(rows, nc * ops.MAX_PER_CLASS)
        self.layout = GatherLayout(global_batch, out_cap)
        self.gatherer = RootGather(self.layout, device, root, group, depth)
        self.pipe = PipelinedDetector(specs, nc, self.last - self.first, device, conf_thres, nms_thres,
                                      depth=depth, use_graph=use_graph, variant=variant)
        for lane, d in enumerate(self.pipe.lanes):
            d.out_ptrs = self.layout.slice_ptrs(self.gatherer.lane_base(lane), self.first)
        self.device = torch.device(device)
        self._host_counts = torch.empty(global_batch, dtype=torch.int32).pin_memory() if self.rank == root else None

    def submit(self, local_heads) -> int:
        return self.pipe.submit(local_heads)

    def wait(self, ticket: int):
        return self.pipe.counts(ticket)[0]        # candidate counts of the local slice (kept counts live on the root)

    def gather(self, ticket: int, return_rows: bool = False, as_list: bool = True):
        """Collective.  Root: the global result of step ``ticket`` -- the reference-shaped list of (n,7) tensors /
        None (``as_list=True``), or the raw ``(out (B,out_cap,7), out_row, counts on the host)`` triple, which skips
        building one Python view per image (tens of milliseconds for thousands of images).  Other ranks: None."""
        self.pipe.lanes[ticket % self.depth].counts()
        dist.barrier(group=self.group)            # every rank's peer stores of this step have completed
        if self.rank != self.root:
            return None
        out, row, cnt = self.gatherer.root_views(ticket % self.depth)
        self._host_counts.copy_(cnt, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        if not as_list:
            return out, row, self._host_counts
        return ops.ragged(out, row, self._host_counts, with_rows=return_rows)

    def close(self):
        self.gatherer.close()
