#!/usr/bin/env python
"""bench.py -- YOLO decode+NMS images/s on B200, with the HBM roofline of the decode/compaction kernel.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME] ...
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

One "step" = one pass of the hot path over one batch of synthetic head tensors: fused decode + confidence
filter + compaction, the three NMS kernels and the read-back of the per-image counts.  Rank 0 prints ONE
JSON line.  Weak scaling: every GPU processes ``--batch`` images per step (global batch = batch * N); kept
rows of all ranks land in rank 0's memory through NVLink peer stores (no collective on the hot path).

The CPU oracle (``oracle/``) is imported only by the ``cpu_baseline`` leg and by ``--impl reference``.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "YOLO decode+NMS images/sec"
UNIT = "images/s"
DEFAULTS = dict(workload="spp-608", batch=64, kind="B", conf=0.3, nms=0.5)
CPU_SAMPLE_BATCH = {"spp-608": 32, "spp-1024": 12, "tiny-416": 128}   # about 10-15 s of CPU work per measurement


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULTS["workload"])
    ap.add_argument("--batch", type=int, default=None, help="images per GPU per step")
    ap.add_argument("--kind", default=DEFAULTS["kind"], help="synthetic input: A (iid logits) or B (+ planted objects)")
    ap.add_argument("--conf", type=float, default=DEFAULTS["conf"])
    ap.add_argument("--nms", type=float, default=DEFAULTS["nms"])
    ap.add_argument("--variant", default="auto", choices=["auto", "ldg", "tma", "tma2d"], help="decode_compact kernel variant")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--depth", type=int, default=None,
                    help="batches in flight (streams): NMS of batch i overlaps decode of i+1; default 6 for batches under 600 MB, else 4")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-head-fusion", action="store_true", help="skip the extra 'head_fusion' measurement (SURVEY 8f-3)")
    ap.add_argument("--cpu-runs", type=int, default=6)
    return ap.parse_args()


def default_batch(workload):
    return {"spp-608": 64, "tiny-416": 1024, "spp-1024": 256}.get(workload, 8)


# ------------------------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index: int, period_s: float = 0.01):
        super().__init__(daemon=True)
        self.index, self.period = index, period_s
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def sample(self):
        nv = self.nv
        self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
        r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
            else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        names = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
                 "hw_power_brake_slowdown": 0x80}
        for k, bit in names.items():
            if r & bit:
                self.reasons.add(k)

    def run(self):
        if not self.ok:
            return
        while not self._halt.is_set():
            try:
                self.sample()
            except Exception:  # noqa: BLE001
                break
            self._halt.wait(self.period)

    def finish(self):
        self._halt.set()
        if self.ok:
            self.join(timeout=2)
            try:
                self.sample()
            except Exception:  # noqa: BLE001
                pass
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def physical_gpu_index(local_index: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_index])
        except Exception:  # noqa: BLE001
            pass
    return local_index


# ------------------------------------------------------------------------------------------- CPU legs
def cpu_oracle_run(workload, batch, kind, conf, nms, runs, warm=1):
    """Times the CPU port of the reference path (oracle/yolo_oracle.py: torch CPU ops of the same granularity
    as the reference) on a bounded sample of the workload.  Returns (images/s median, per-run seconds)."""
    from oracle import yolo_oracle
    from pytorch_yolo_b200 import synth
    w = synth.WORKLOADS[workload]
    heads = synth.synth_heads(workload, batch, kind, seed=1234)
    times = []
    with torch.no_grad():
        for i in range(warm + runs):
            t0 = time.perf_counter()
            pred = yolo_oracle.decode_heads(heads, w["anchors"], w["nc"], w["img_size"])
            yolo_oracle.non_max_suppression(pred, conf, nms)
            dt = time.perf_counter() - t0
            if i >= warm:
                times.append(dt)
    return batch / statistics.median(times), times


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (the oracle port: the Python reference
    itself cannot travel to the GPU box) on the host cores, same metric / unit / config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sample_b = CPU_SAMPLE_BATCH.get(args.workload, 8)
    steps = max(1, min(args.steps, 8))
    warm = max(1, min(args.warmup, 2))
    ips, times = cpu_oracle_run(args.workload, sample_b, args.kind, args.conf, args.nms, runs=steps, warm=warm)
    line = {
        "metric": METRIC, "value": ips, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": 1e3 * statistics.median(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, sample_b, cpu=True),
        "cpu_baseline": {"value": ips, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{args.workload} SYNTH-{args.kind} batch {sample_b} per step, conf {args.conf} nms {args.nms}"},
        "e2e": {"value": ips, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, batch, cpu=False):
    from pytorch_yolo_b200 import synth
    return {"workload": f"{args.workload} ({synth.anchors_per_image(args.workload)} anchors/img, 80 classes) "
                        f"SYNTH-{args.kind} heads, batch {batch}{'' if cpu else '/GPU'}, conf {args.conf} nms {args.nms}",
            "batch_per_gpu": batch, "conf_thres": args.conf, "nms_thres": args.nms,
            "l2": "inputs exceed L2 (no flush needed)" if synth.head_bytes_per_image(args.workload) * batch > 126e6
                  else "inputs rotate through >L2 worth of buffers",
            "sharding": "images across GPUs, kept rows to rank 0 by NVLink peer stores"}


# ------------------------------------------------------------------------------------------- head fusion
def head_fusion_probe(args, w, specs, B, dev, peak_gbs):
    """Extra measurement, not part of `value`: the head 1x1 convolutions fused with decode + compaction on the tensor
    cores (csrc/head.cu) against the unfused sequence -- torch's cuDNN convolution (TF32) writing the head tensors, then
    decode_compact reading them -- on synthetic feature maps of the workload's shapes.  CUDA events, 20 repetitions."""
    from pytorch_yolo_b200 import ops, synth
    feats, convs = synth.synth_head_convs(args.workload, B, device=dev)
    nc = w["nc"]
    rows = sum(s.rows for s in specs)
    offs, o = [], 0
    for s in specs:
        offs.append(o)
        o += s.rows
    hws = [ops.fold_head(c, dev) for c in convs]
    buf = ops.Buffers(dev, B, rows, nc)
    padded = [None if ops.head_supported(h.c_in, s, nc) else torch.zeros(B, h.c_in, ops.padded_pitch(s), device=dev)
              for h, s in zip(hws, specs)]

    def fused():
        xs = [x if p is None else ops.pad_feature(x, out=p) for x, p in zip(feats, padded)]
        ops.head_decode_compact(xs, hws, specs, offs, rows, nc, args.conf, buf)

    def timeit(fn, reps=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / reps * 1e3

    t_fused = timeit(fused)
    cand = int(buf.meta[:B].sum())
    ovf = int(buf.meta[B])
    with torch.no_grad():
        t_conv = timeit(lambda: [c(x) for c, x in zip(convs, feats)])
        heads = [c(x) for c, x in zip(convs, feats)]
        t_dec = timeit(lambda: ops.decode_compact(heads, specs, nc, args.conf, buf))
    x_bytes = sum(x.numel() * 4 for x in feats)
    flops = sum(2.0 * B * s.ny * s.nx * h.n_out * h.c_in for s, h in zip(specs, hws))
    return {"what": "1x1 head conv (TF32 tcgen05) + decode + compaction in one kernel, all scales in one launch; "
                    "planes that are not a multiple of 4 floats go through a padded copy first (included)",
            "fused_us": t_fused, "unfused_us": t_conv + t_dec, "unfused_conv_cudnn_us": t_conv, "unfused_decode_compact_us": t_dec,
            "speedup": (t_conv + t_dec) / t_fused, "feature_bytes": x_bytes, "feature_gbs": x_bytes / t_fused / 1e3,
            "frac_of_hbm_peak": x_bytes / t_fused / 1e3 / peak_gbs, "tf32_tflops": flops / t_fused / 1e6,
            "candidates": cand, "overflow": ovf, "padded_scales": [p is not None for p in padded],
            "allow_tf32_reference": bool(torch.backends.cudnn.allow_tf32)}


# ------------------------------------------------------------------------------------------- GPU arm
def main():
    args = parse_args()
    if args.batch is None:
        args.batch = default_batch(args.workload)
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    from pytorch_yolo_b200 import ops, synth
    from pytorch_yolo_b200.detect import PipelinedDetector
    from pytorch_yolo_b200.sharded import ShardedDetector

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1 and args.gpus > 1:
        raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    distributed = world > 1
    if distributed:
        dist.init_process_group("nccl", device_id=dev)

    w = synth.WORKLOADS[args.workload]
    B = args.batch
    # deeper pipelines hide the fixed per-batch latencies of small batches; big batches gain nothing and lose cache
    depth = max(1, args.depth) if args.depth is not None else (6 if synth.head_bytes_per_image(args.workload) * B < 600e6 else 4)
    specs = [ops.scale_spec(a, g, g, w["img_size"]) for a, g in zip(w["anchors"], w["grids"])]
    bytes_per_img = synth.head_bytes_per_image(args.workload)

    # inputs: resident in HBM; when one batch is smaller than L2, rotate through enough distinct batches
    n_sets = min(64, max(1, int(-(-160e6 // (bytes_per_img * B)))))
    head_sets = [synth.synth_heads(args.workload, B, args.kind, seed=1234 + 7919 * rank + s, device=dev)
                 for s in range(n_sets)]
    torch.cuda.synchronize(dev)

    use_graph = not args.no_graph
    if distributed:
        # one sharded pipeline; with several input sets the pointers change per step -> eager launches
        det = ShardedDetector(specs, w["nc"], B * world, dev, args.conf, args.nms,
                              use_graph=use_graph and n_sets == 1, depth=depth, variant=args.variant)
        pipes = [det.pipe] * n_sets
    else:
        det = None
        # one pipeline (one captured graph per lane) per input set so that graph replay sees static pointers
        pipes = [PipelinedDetector(specs, w["nc"], B, dev, args.conf, args.nms, depth=depth, use_graph=use_graph,
                                   variant=args.variant) for _ in range(n_sets)]
    lane0 = pipes[0].lanes[0]

    def run_steps(k):
        """k steps with `depth` batches in flight: submit step i, then wait for step i - depth + 1."""
        pending, last = [], None
        for i in range(k):
            p = pipes[i % n_sets]
            pending.append((p, p.submit(head_sets[i % n_sets])))
            if len(pending) >= depth:
                q, t = pending.pop(0)
                last = q.counts(t)[0]
        for q, t in pending:
            last = q.counts(t)[0]
        return last

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize(dev)

    cand_total = int(run_steps(max(args.warmup, 3)).sum())

    sampler = ClockSampler(physical_gpu_index(local_rank))
    stream = torch.cuda.current_stream(dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.start()
    ev0.record(stream)
    run_steps(args.steps)                    # every step's counts were read back (host sync per step)
    if distributed:
        det.gather(det.pipe._next - 1, as_list=False)   # barrier + the root reads the gathered counts: ragged gather complete
    for p in set(pipes):
        p.drain()                            # the timing stream waits for the pipeline streams
    ev1.record(stream)
    barrier()
    clocks = sampler.finish()
    elapsed_ms = ev0.elapsed_time(ev1)
    by_rank = [elapsed_ms / args.steps]
    if distributed:
        t = torch.tensor([elapsed_ms], device=dev)
        every = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(every, t)
        by_rank = [float(x.item()) / args.steps for x in every]
        elapsed_ms = max(float(x.item()) for x in every)          # the job is as slow as its slowest rank
    value = B * world * args.steps / (elapsed_ms * 1e-3)

    # ---- roofline of the dominant kernel (decode_compact), timed alone with CUDA events on its stream
    buf = lane0.buf
    reps = max(20, min(args.steps, 200))
    for s in range(min(3, n_sets)):
        ops.decode_compact(head_sets[s], specs, w["nc"], args.conf, buf, variant=args.variant)
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    k0.record(stream)
    for i in range(reps):
        ops.decode_compact(head_sets[i % n_sets], specs, w["nc"], args.conf, buf, variant=args.variant)
    k1.record(stream)
    torch.cuda.synchronize(dev)
    kern_ms = k0.elapsed_time(k1) / reps
    algo_bytes = B * bytes_per_img + 32 * cand_total
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = algo_bytes / (kern_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.isfile(tpath):
        try:
            traffic = json.load(open(tpath)).get(f"{args.workload}:{B}:{args.conf}")
        except Exception:  # noqa: BLE001
            traffic = None
    roofline = {"bound": "hbm", "kernel": "decode_compact_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "variant": args.variant, "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": algo_bytes, "kernel_ms": kern_ms, "candidates_per_launch": cand_total}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(args, B), "clocks": clocks,
        "gpu_launches": lane0.kernels_per_step * args.steps, "roofline": roofline,
        "cuda_graph": bool(lane0.use_graph), "batches_in_flight": depth, "ms_per_step_by_rank": by_rank,
    }

    # ---- the next row of the scope table (8f-3), measured beside the headline: head convolution fused in
    if rank == 0 and world == 1 and not args.no_head_fusion and "head_cin" in w:
        try:
            line["head_fusion"] = head_fusion_probe(args, w, specs, B, dev, peak)
        except Exception as e:  # noqa: BLE001  (the headline must not depend on this extra)
            line["head_fusion"] = {"error": repr(e)[:300]}

    # ---- end to end through the public API with HOST buffers (pinned), H2D + D2H inside the timed region
    if not args.no_e2e:
        host_heads = [h.cpu().pin_memory() for h in head_sets[0]]
        host_out = torch.empty(B, lane0.buf.out_cap, ops.DET_COLS, dtype=torch.float32).pin_memory()
        e2e_steps = max(3, min(args.steps, 30))

        def e2e_step():
            if not distributed:
                return lane0.run_from_host(host_heads, head_sets[0], host_out)
            h2d = 0
            for dst, src in zip(head_sets[0], host_heads):
                dst.copy_(src, non_blocking=True)
                h2d += src.numel() * 4
            det.wait(det.submit(head_sets[0]))
            return None, None, h2d, lane0.buf.meta_host.numel() * 4

        for _ in range(3):
            e2e_step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(e2e_steps):
            kept, _, h2d, d2h = e2e_step()
        if distributed:
            last = det.pipe._next - 1
            res = det.gather(last, as_list=False)   # root: all ranks' kept rows are now in its memory
            if rank == 0:
                n_max = max(1, int(res[2].max()))
                out_all = res[0]
                host_all = torch.empty(B * world, n_max, ops.DET_COLS, dtype=torch.float32).pin_memory()
                host_all.copy_(out_all[:, :n_max], non_blocking=True)
                d2h += host_all.numel() * 4 // e2e_steps
            det.pipe.drain()
        e1.record(stream)
        barrier()
        e2e_ms = e0.elapsed_time(e1)
        if distributed:
            t = torch.tensor([e2e_ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_ms = float(t.item())
        line["e2e"] = {"value": B * world * e2e_steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                       "d2h_bytes_per_step": d2h, "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
                       "kept_rows_per_step": int(kept.sum()) if kept is not None else None,
                       "path": "pinned host heads -> H2D -> Detector (C ABI) -> D2H counts + kept rows"}

    # ---- CPU baseline beside it (rank 0, N=1 only): the oracle port on a bounded sample
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        sb = CPU_SAMPLE_BATCH.get(args.workload, 8)
        ips, times = cpu_oracle_run(args.workload, sb, args.kind, args.conf, args.nms, runs=args.cpu_runs)
        line["cpu_baseline"] = {"value": ips, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{args.workload} SYNTH-{args.kind} batch {sb}, {args.cpu_runs} runs (median), "
                                          f"{sum(times):.1f} s of CPU work"}

    if rank == 0:
        print(json.dumps(line), flush=True)
    if distributed:
        det.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
