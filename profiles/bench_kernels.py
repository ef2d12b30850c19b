"""Per-kernel timing of both paths (fused and API/dense) for one workload, CUDA events, inputs resident.
    python profiles/bench_kernels.py spp-608 64 0.3 [kind]
Prints the algorithmic bytes and achieved GB/s of the three streaming kernels and the NMS stage time."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pytorch_yolo_b200 import ops, synth  # noqa: E402


def timeit(fn, reps=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3   # us


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "spp-608"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    conf = float(sys.argv[3]) if len(sys.argv) > 3 else 0.3
    kind = sys.argv[4] if len(sys.argv) > 4 else "B"
    dev = "cuda:0"
    w = synth.WORKLOADS[wl]
    n = synth.anchors_per_image(wl)
    heads = synth.synth_heads(wl, B, kind, seed=1234, device=dev)
    specs = [ops.scale_spec(a, g, g, w["img_size"]) for a, g in zip(w["anchors"], w["grids"])]
    buf = ops.Buffers(dev, B, n, w["nc"])
    out, out_row = buf.new_outputs()
    T = B * synth.head_bytes_per_image(wl)
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.isfile(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0

    ops.decode_compact(heads, specs, w["nc"], conf, buf)
    cand, _, _ = ops.read_counts(buf)
    K = int(cand.sum())
    ops.nms(buf, 0.5, out, out_row)
    _, kept, _ = ops.read_counts(buf)
    res = {"workload": wl, "batch": B, "conf": conf, "kind": kind, "head_bytes": T, "candidates": K,
           "kept": int(kept.sum()), "max_cand_per_img": int(cand.max())}

    t = timeit(lambda: ops.decode_compact(heads, specs, w["nc"], conf, buf, variant="ldg"))
    res["decode_compact_us"] = t
    res["decode_compact_gbs"] = (T + 32 * K) / t / 1e3
    t = timeit(lambda: ops.nms(buf, 0.5, out, out_row))
    res["nms_3kernels_us"] = t
    pred = torch.empty(B, n, w["nc"] + 5, device=dev)
    t = timeit(lambda: ops.decode_dense(heads, specs, w["nc"], out=pred))
    res["decode_dense_us"] = t
    res["decode_dense_gbs"] = 2 * T / t / 1e3
    t = timeit(lambda: ops.compact_from_dense(pred, conf, buf, write_back=False))
    res["compact_from_dense_us"] = t
    res["compact_from_dense_gbs"] = (T + 32 * K) / t / 1e3
    for k in ("decode_compact", "decode_dense", "compact_from_dense"):
        res[k + "_frac_of_measured_peak"] = res[k + "_gbs"] / peak
    print(json.dumps(res))


if __name__ == "__main__":
    main()
