"""Multi-GPU: image batches shard across ranks, kept rows gather on the root over NVLink peer stores.

The reference has no multi-GPU path.  Images are independent, so the path shards
with no data-path collective: one process per GPU (``torch.distributed`` is used only to exchange a
64-byte CUDA-IPC handle at set-up), every rank runs decode_compact + NMS on its own contiguous slice of the
batch, and its ``nms_finalize_kernel`` writes the kept rows of image ``g`` straight into ``root_out[g]`` /
``root_count[g]`` -- the root GPU's memory, mapped through CUDA IPC -- so the "ragged gather" is a handful of
coalesced peer stores per image instead of a collective.

Step protocol (no NCCL call and no host round trip per batch; everything below is enqueued work, replayable from a
CUDA graph).  The root buffer has ``depth`` lanes; lane ``l`` carries, next to its rows, one completion stamp per rank
and one ``ack`` word:

* every rank's finalize kernel publishes ``stamp[l][rank] = u + 1`` (release, system scope) once all rows of its
  ``u``-th use of lane ``l`` are stored (``yolo_b200_nms_ex``);
* the root waits, on the GPU, for all ``world`` stamps of the lane (``yolo_b200_flag_wait``) before it copies the
  global counts to the host: after ``gather`` the rows of every rank are in the root's memory;
* a lane is re-used ``depth`` steps later.  The root begins its own use ``u`` of the lane by publishing
  ``ack[l] = u`` (``yolo_b200_flag_post``, ordered after everything the caller enqueued before ``submit``); every other
  rank begins use ``u`` by waiting for it.  So no rank overwrites a lane before the root has started the matching step,
  i.e. the results ``gather(t)`` returned stay valid until the root calls ``submit(t + depth)``.
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
    """Byte layout of one lane of the root's result buffer:
    out (B, out_cap, 7) f32 | out_row (B, out_cap) i32 | out_count (B) i32 | stamp (world) i32 | ack (1) i32."""
    global_batch: int
    out_cap: int
    world: int = 1

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
    def stamp_off(self) -> int:
        return self.count_off + _align(self.global_batch * 4)

    @property
    def ack_off(self) -> int:
        return self.stamp_off + 4 * self.world

    @property
    def total(self) -> int:
        return self.stamp_off + _align(4 * (self.world + 1))

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
        self._views = {}
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
        """(out, out_row, out_count) tensors over copy ``lane`` of the root buffer -- root rank only (built once per lane:
        a dozen tensor operations per call would sit in the per-step host loop of the gather)."""
        cached = self._views.get(lane)
        if cached is None:
            cached = self._views[lane] = self._make_views(lane)
        return cached

    def _make_views(self, lane: int):
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
                self._views.clear()
                self.bytes = None
                lib.yolo_b200_device_free(C.c_void_p(self._owned)); self._owned = None


class ShardedDetector:
    """One per rank.  ``submit(local_inputs)`` runs the fused path on this rank's image slice (pipelined over ``depth``
    streams) and stores its kept rows into the root's buffer; ``gather(ticket)`` waits for that step -- on the root for
    the rows of EVERY rank (device-side step flags, see the module docstring) -- and returns the global ragged list on
    the root and ``None`` elsewhere.  Contract: all ranks issue the same sequence of ``submit`` / ``gather`` calls, and
    the root has finished with the result of ``gather(t)`` when it calls ``submit(t + depth)``.

    ``heads``: head modules (1x1 convolutions) -- the lanes become :class:`pytorch_yolo_b200.head.HeadDetector` and
    ``submit`` takes this rank's feature maps instead of head tensors."""

    def __init__(self, specs: Sequence[ops.ScaleSpec], nc: int, global_batch: int, device,
                 conf_thres: float = 0.5, nms_thres: float = 0.5, root: int = 0, group=None,
                 use_graph: bool = True, depth: int = 2, variant: str = "auto", heads=None, timeout_s: float = 20.0):
        from .detect import Detector, PipelinedDetector
        self.group, self.root, self.depth = group, root, depth
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.is_root = self.rank == root
        self.plan = ShardPlan(global_batch, self.world)
        self.first, self.last = self.plan.bounds(self.rank)
        self.device = torch.device(device)
        rows = sum(s.rows for s in specs)
        out_cap = min(rows, nc * ops.MAX_PER_CLASS)
        L = self.layout = GatherLayout(global_batch, out_cap, self.world)
        self.gatherer = RootGather(self.layout, device, root, group, depth)
        # device-side use counters of this rank's flag kernels, one set per lane: [step | ack]
        self._seq = torch.zeros(depth, 2, dtype=torch.int32, device=self.device)
        self._host_counts = [torch.zeros(global_batch, dtype=torch.int32).pin_memory() for _ in range(depth)] \
            if self.is_root else None
        local = self.last - self.first
        dev = self.device

        def make(lane: int):
            base = self.gatherer.lane_base(lane)
            seq = self._seq[lane].data_ptr()
            stamps, ack = base + L.stamp_off, base + L.ack_off
            kw = dict(use_graph=use_graph, out_ptrs=L.slice_ptrs(base, self.first), nms_priority=True,
                      step=(seq, stamps + 4 * self.rank))
            det = (Detector(specs, nc, local, dev, conf_thres, nms_thres, variant=variant, **kw) if heads is None else
                   _head_detector(heads, specs, nc, local, dev, conf_thres, nms_thres, **kw))
            err = det.buf.sync_err_ptr
            if self.is_root:
                wait_seq = torch.zeros(1, dtype=torch.int32, device=dev)
                cnt = self.gatherer.root_views(lane)[2]
                host = self._host_counts[lane]
                det._keep = (wait_seq, cnt)
                det.pre_hook = lambda: ops.flag_post(ack, seq + 4, -1, dev)                       # ack[lane] = u
                det.post_hook = lambda: (ops.flag_wait(stamps, self.world, wait_seq.data_ptr(), 0, err, dev, timeout_s),
                                         host.copy_(cnt, non_blocking=True))                       # every stamp >= u + 1
            else:
                det.pre_hook = lambda: ops.flag_wait(ack, 1, seq + 4, -1, err, dev, timeout_s)     # ack[lane] >= u
            det.kernels_per_step += 2 if self.is_root else 1
            return det

        self.pipe = PipelinedDetector(specs, nc, local, device, conf_thres, nms_thres, depth=depth, factory=make)

    def bind(self, local_inputs, per_lane: bool = False) -> None:
        """Collective: capture every lane's graph for these static inputs (runs each lane once on every rank)."""
        self.pipe.bind(local_inputs, per_lane=per_lane)

    def submit(self, local_inputs) -> int:
        return self.pipe.submit(local_inputs)

    def wait(self, ticket: int):
        return self.pipe.counts(ticket)[0]        # candidate counts of the local slice (kept counts live on the root)

    def gather(self, ticket: int, return_rows: bool = False, as_list: bool = True):
        """Root: the global result of step ``ticket`` -- the reference-shaped list of (n,7) tensors / None
        (``as_list=True``), or the raw ``(out (B,out_cap,7), out_row, counts on the host)`` triple, which skips building
        one Python view per image (tens of milliseconds for thousands of images).  Other ranks: None, once their own part
        of the step is done.  No collective call: the root's lane stream already waited for every rank's stamp."""
        lane = ticket % self.depth
        self.pipe.lanes[lane].counts()            # host waits for this rank's lane stream; raises on overflow / time-out
        if not self.is_root:
            return None
        out, row, _ = self.gatherer.root_views(lane)
        counts = self._host_counts[lane]
        if not as_list:
            return out, row, counts
        return ops.ragged(out, row, counts, with_rows=return_rows)

    def close(self):
        self.gatherer.close()


def _head_detector(heads, *args, **kw):
    from .head import HeadDetector
    return HeadDetector(heads, *args, **kw)
