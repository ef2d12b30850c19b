"""CPU (gloo, world size 2): host-side logic of the multi-GPU path -- shard bounds, gather layout, the IPC-handle
exchange -- with the per-rank compute replaced by the oracle (tests only).  The peer-store data path itself needs
GPUs (tests/multi_gpu_check.py, run under torchrun on >= 2 GPUs)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pytorch_yolo_b200 import ops, sharded, synth


def test_shard_plan_covers_batch_exactly_once():
    for batch, world in [(64, 8), (1024, 8), (10, 4), (3, 8), (7, 2), (256, 1)]:
        plan = sharded.ShardPlan(batch, world)
        seen = []
        for r in range(world):
            s, e = plan.bounds(r)
            assert 0 <= s <= e <= batch and e - s == plan.local_batch(r)
            seen += list(range(s, e))
            for g in range(s, e):
                assert plan.owner(g) == r
        assert seen == list(range(batch))
        sizes = [plan.local_batch(r) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1


def test_gather_layout_slices_do_not_overlap():
    lay = sharded.GatherLayout(global_batch=16, out_cap=100)
    assert lay.row_off % 256 == 0 and lay.count_off % 256 == 0 and lay.total % 256 == 0
    assert lay.row_off >= 16 * 100 * 7 * 4 and lay.count_off - lay.row_off >= 16 * 100 * 4
    base = 1 << 20
    o0, r0, c0 = lay.slice_ptrs(base, 0)
    o1, r1, c1 = lay.slice_ptrs(base, 8)
    assert (o1 - o0, r1 - r0, c1 - c0) == (8 * 100 * 28, 8 * 100 * 4, 8 * 4)
    assert o0 == base and r0 == base + lay.row_off and c0 == base + lay.count_off
    # step flags behind the counts: one completion stamp per rank, one ack word, inside the lane
    lay8 = sharded.GatherLayout(global_batch=16, out_cap=100, world=8)
    assert lay8.stamp_off >= lay8.count_off + 16 * 4 and lay8.stamp_off % 256 == 0
    assert lay8.ack_off == lay8.stamp_off + 8 * 4 and lay8.total >= lay8.ack_off + 4 and lay8.total % 256 == 0
    assert lay8.slice_ptrs(base, 8) == lay.slice_ptrs(base, 8)          # rows / counts do not move with the world size


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1. the 64-byte handle travels from the root to every rank
        handle = bytes(range(64)) if rank == 0 else None
        got = sharded.exchange_handle(handle, root=0)
        assert got == bytes(range(64))
        # 2. each rank processes its slice (oracle stands in for the kernels) and "stores" into the root layout
        from oracle import yolo_oracle
        batch = 5
        plan = sharded.ShardPlan(batch, world)
        s, e = plan.bounds(rank)
        pred = synth.synth_prediction(batch, 300, nc=6, seed=7)
        local = yolo_oracle.non_max_suppression(pred[s:e].clone(), 0.2, 0.5)
        out_cap = 300
        lay = sharded.GatherLayout(batch, out_cap)
        out = torch.zeros(batch, out_cap, 7)
        cnt = torch.zeros(batch, dtype=torch.int32)
        for k, d in enumerate(local):
            if d is not None:
                out[s + k, :len(d)] = d
                cnt[s + k] = len(d)
        dist.all_reduce(out)         # disjoint slices: the sum is the gather (gloo stands in for NVLink peer stores)
        dist.all_reduce(cnt)
        if rank == 0:
            want = yolo_oracle.non_max_suppression(pred.clone(), 0.2, 0.5)
            got_list = ops.ragged(out, None, cnt)
            for g, w_ in zip(got_list, want):
                assert (g is None) == (w_ is None)
                if g is not None:
                    assert torch.equal(g, w_)
            assert lay.total > 0
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_two_rank_gather_over_gloo():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        assert dict(ret) == {0: "ok", 1: "ok"}


def test_c_abi_exports_every_declared_symbol():
    """The library loads here (no GPU) and exports exactly what include/yolo_b200.h declares."""
    import re
    from pytorch_yolo_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "yolo_b200.h")).read()
    declared = set(re.findall(r"\b(yolo_b200_[a-z_0-9]+)\s*\(", header))
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.exported_symbols())
    assert lib.yolo_b200_abi_version() == _lib.ABI_VERSION
    # argument validation happens before any CUDA call
    assert lib.yolo_b200_nms(None, None, None, 1, 1, 1, 0.5, 100, None, None, 1, None, None, 0, None) == -1
    assert lib.yolo_b200_nms_workspace_bytes(0, 0, 0, 0) == 0
    assert lib.yolo_b200_nms_workspace_bytes(64, 22743, 80, 100) > 64 * 22743 * 12


def test_product_does_not_import_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "pytorch_yolo_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                imports = [l for l in src.splitlines() if l.strip().startswith(("import ", "from "))]
                assert not any("oracle" in l for l in imports), f"{f} imports the oracle"


def test_cpu_tensors_fail_loudly():
    from pytorch_yolo_b200 import YOLOLayer, non_max_suppression
    with pytest.raises(ops.YoloB200Error):
        non_max_suppression(torch.zeros(1, 10, 85), 0.3, 0.5)
    layer = YOLOLayer(((10.0, 13.0),), 80, [((10.0, 13.0),)]).eval()
    with pytest.raises(ops.YoloB200Error):
        layer(torch.zeros(1, 85, 4, 4), 128)
    # training mode is a pure reshape and needs no kernel (reference yolo_layer.py:67-72)
    layer.train()
    assert layer(torch.zeros(1, 85, 4, 4), 128).shape == (1, 1, 4, 4, 85)


def test_ragged_views_are_the_per_image_slices():
    """ops.ragged: the one-call split gives exactly the views a slice per image gives (same memory, shape, strides, writable),
    None for images without detections, and the non-contiguous fallback agrees."""
    g = torch.Generator().manual_seed(3)
    out = torch.rand(7, 50, 7, generator=g)
    row = torch.randint(0, 1000, (7, 50), generator=g, dtype=torch.int32)
    cnt = torch.tensor([3, 0, 50, 1, 0, 17, 49], dtype=torch.int32)
    dets, rows = ops.ragged(out, row, cnt, with_rows=True)
    assert ops.ragged(out, None, cnt)[2].data_ptr() == out[2].data_ptr()
    for i, n in enumerate(cnt.tolist()):
        if n == 0:
            assert dets[i] is None and rows[i] is None
            continue
        for got, want in ((dets[i], out[i, :n]), (rows[i], row[i, :n])):
            assert got.data_ptr() == want.data_ptr() and got.shape == want.shape and got.stride() == want.stride()
            assert torch.equal(got, want)
    dets[0][0, 0] = -5.0                         # the caller may write into what it got (reference utils.py:313)
    assert out[0, 0, 0] == -5.0
    wide = torch.rand(7, 50, 9, generator=g)[:, :, :7]          # not contiguous: the per-image path
    d2 = ops.ragged(wide, None, cnt)
    assert d2[1] is None and torch.equal(d2[5], wide[5, :17]) and d2[5].data_ptr() == wide[5].data_ptr()
