"""Multi-GPU parity check, launched by torchrun (one rank per GPU):

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/multi_gpu_check.py

Every rank runs the fused path on its slice of a seeded global batch; the kept rows land in rank 0's memory
through NVLink peer stores; rank 0 compares the gathered ragged result with the single-GPU result for the whole
batch -- bit for bit, anchor rows included."""
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
    failures = 0
    for wl, batch, conf in (("tiny-416", 10, 0.3), ("spp-608", 5, 0.3), ("mini-160", 7, 0.01)):
        w = synth.WORKLOADS[wl]
        specs = [ops.scale_spec(a, g, g, w["img_size"]) for a, g in zip(w["anchors"], w["grids"])]
        heads = synth.synth_heads(wl, batch, "B", seed=4321)              # same global batch on every rank (CPU)
        det = ShardedDetector(specs, w["nc"], batch, dev, conf, 0.5, depth=2)
        local_heads = [h[det.first:det.last].contiguous().to(dev) for h in heads]
        for _ in range(3):                                                 # several steps through both lanes
            t = det.submit(local_heads)
            det.wait(t)
            res = det.gather(t, return_rows=True)
        if rank == 0:
            single = Detector(specs, w["nc"], batch, dev, conf, 0.5, use_graph=False)
            want, want_rows = single.run([h.to(dev) for h in heads], return_rows=True, clone=True)
            got, got_rows = res
            for i, (g, gr, o, orow) in enumerate(zip(got, got_rows, want, want_rows)):
                same = (g is None) == (o is None) and (g is None or (torch.equal(g, o) and torch.equal(gr, orow)))
                if not same:
                    failures += 1
                    print(f"MISMATCH {wl} image {i}")
        det.close()
    flag = torch.tensor([failures], device=dev)
    dist.all_reduce(flag)
    if rank == 0:
        print("multi-gpu check:", "FAILED" if int(flag) else f"ok ({world} ranks)")
    dist.destroy_process_group()
    sys.exit(1 if int(flag) else 0)


if __name__ == "__main__":
    main()
