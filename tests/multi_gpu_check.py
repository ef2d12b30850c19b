"""Multi-GPU parity check, launched by torchrun (one rank per GPU):

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/multi_gpu_check.py

Every rank runs the fused path on its slice of a seeded global batch; the kept rows land in rank 0's memory
through NVLink peer stores and rank 0 learns about their completion from the device-side step stamps (no collective);
rank 0 compares the gathered ragged result with the single-GPU result for the whole batch -- bit for bit, anchor rows
included.  The input CHANGES every step and several steps are in flight (lanes are re-used), so a result assembled
from two different steps, or a lane overwritten before the root read it, shows up as a mismatch."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pytorch_yolo_b200 import ops, synth                     # noqa: E402
from pytorch_yolo_b200.detect import Detector                # noqa: E402
from pytorch_yolo_b200.sharded import ShardedDetector        # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    failures = checked = 0
    for wl, batch, conf, depth, graph in (("tiny-416", 10, 0.3, 2, True), ("spp-608", 5, 0.3, 3, True),
                                          ("mini-160", 7, 0.01, 2, False), ("tiny-416", 2 * world + 1, 0.05, 4, True),
                                          ("mini-160", world - 1, 0.01, 2, True)):       # the last rank has no image at all
        w = synth.WORKLOADS[wl]
        specs = [ops.scale_spec(a, g, g, w["img_size"]) for a, g in zip(w["anchors"], w["grids"])]
        n_steps = 3 * depth + 1
        # a different global batch per step, identical on every rank (CPU generator)
        batches = [synth.synth_heads(wl, batch, "B", seed=4321 + 17 * s) for s in range(n_steps)]
        det = ShardedDetector(specs, w["nc"], batch, dev, conf, 0.5, depth=depth, use_graph=graph)
        local_sets = [[h[det.first:det.last].contiguous().to(dev) for h in b] for b in batches]
        # graph replay needs static input tensors: one set per lane, refilled before the lane is re-used
        statics = [[t.clone() for t in local_sets[0]] for _ in range(depth)]
        if graph:
            det.bind(statics, per_lane=True)
        single = Detector(specs, w["nc"], batch, dev, conf, 0.5, use_graph=False) if rank == 0 else None
        pending = []

        def check(step, res):
            nonlocal failures, checked
            if rank != 0:
                return
            want, want_rows = single.run([h.to(dev) for h in batches[step]], return_rows=True, clone=True)
            got, got_rows = res
            for i, (g, gr, o, orow) in enumerate(zip(got, got_rows, want, want_rows)):
                same = (g is None) == (o is None) and (g is None or (torch.equal(g, o) and torch.equal(gr, orow)))
                checked += 1
                if not same:
                    failures += 1
                    print(f"MISMATCH {wl} step {step} image {i}")

        for s in range(n_steps):                  # `depth` steps in flight, every one on different data
            lane_in = statics[s % depth]
            for d, src in zip(lane_in, local_sets[s]):
                d.copy_(src)                      # ordered before the launch: submit() waits for the caller's stream
            pending.append((s, det.submit(lane_in)))
            if len(pending) >= depth:
                s0, t0 = pending.pop(0)
                check(s0, det.gather(t0, return_rows=True))
        for s0, t0 in pending:
            check(s0, det.gather(t0, return_rows=True))
        det.close()
    # the same protocol with the head convolution fused in: ShardedDetector over HeadDetector lanes (SURVEY 8f-3 + 8e)
    from pytorch_yolo_b200.head import HeadDetector          # noqa: E402
    wl, batch, conf, depth = "tiny-416", 2 * world + 1, 0.3, 2
    w = synth.WORKLOADS[wl]
    specs = [ops.scale_spec(a, g, g, w["img_size"]) for a, g in zip(w["anchors"], w["grids"])]
    feats, convs = synth.synth_head_convs(wl, batch, seed=77, device=dev)        # identical on every rank (seeded per device type)
    det = ShardedDetector(specs, w["nc"], batch, dev, conf, 0.5, depth=depth, heads=convs, use_graph=False)
    local = [f[det.first:det.last].contiguous() for f in feats]
    tickets = [det.submit(local) for _ in range(depth)]
    results = [det.gather(t, return_rows=True) for t in tickets]
    if rank == 0:
        want, want_rows = HeadDetector(convs, specs, w["nc"], batch, dev, conf, 0.5).run(feats, return_rows=True, clone=True)
        for got, got_rows in results:
            for i, (g, gr, o, orow) in enumerate(zip(got, got_rows, want, want_rows)):
                same = (g is None) == (o is None) and (g is None or (torch.equal(g, o) and torch.equal(gr, orow)))
                checked += 1
                if not same:
                    failures += 1
                    print(f"MISMATCH fused head image {i}")
    det.close()
    flag = torch.tensor([failures], device=dev)
    dist.all_reduce(flag)
    if rank == 0:
        print("multi-gpu check:", "FAILED" if int(flag) else f"ok ({world} ranks, {checked} image results compared)")
    dist.destroy_process_group()
    sys.exit(1 if int(flag) else 0)


if __name__ == "__main__":
    main()
