"""Table of the driver-command runs at N = 1, 2, 4, 8 (bench.py --gpus N --steps 20 --warmup 5):
    python profiles/summarize_scaling.py gpurun_out/r02z_bench_n1.json gpurun_out/r02s_bench_n2.json ... > profiles/r02_h_scaling.txt"""
import json
import sys


def last_json(path):
    line = [l for l in open(path) if l.startswith("{")][-1]
    return json.loads(line)


def main(paths):
    rows = [last_json(p) for p in paths]
    rows.sort(key=lambda d: d["n_gpus"])
    base = rows[0]["value"] / rows[0]["n_gpus"]
    print("# bench.py --gpus N --steps 20 --warmup 5 (torchrun, one rank per GPU), spp-608 batch 64 per GPU, conf 0.3 / nms 0.5")
    print("# value = images/s over all ranks (max over ranks of the device time); efficiency = value / (N x value at N = 1)")
    print(f"{'N':>2} {'images/s':>11} {'ms/step':>8} {'eff':>6} {'ms/step by rank':<58} {'warm':>5} {'gather':>7} {'e2e img/s':>10}")
    for d in rows:
        n = d["n_gpus"]
        by = " ".join(f"{x * 1e3:.1f}" for x in d["ms_per_step_by_rank"])
        print(f"{n:>2} {d['value']:>11.0f} {d['ms_per_step']:>8.4f} {d['value'] / (n * base):>6.3f} {by:<58} {d['warmup']:>5} "
              f"{str(d.get('gather_check', '-')):>7} {d.get('e2e', {}).get('value', float('nan')):>10.0f}")
    print()
    print("# other BASELINE configs in the same records (images/s, ms/step, fraction of the decode-traffic HBM floor)")
    for d in rows:
        for c in d.get("configs", []):
            if "value" in c:
                print(f"N={d['n_gpus']}  {c['config']:<42} {c['value']:>11.0f} {c['ms_per_step'] * 1e3:>8.1f} us {c['step_floor_frac']:>6.3f} "
                      f"{str(c.get('gather_check', '-')):>4}")
    print()
    print("# fused-head step (feature maps -> kept detections, sharded like the headline path): head_fusion.pipeline")
    for d in rows:
        p = d.get("head_fusion", {}).get("pipeline", {})
        if "ms_per_step" in p:
            print(f"N={d['n_gpus']}  {p['images_per_s']:>11.0f} images/s  {p['ms_per_step'] * 1e3:>7.1f} us/step  gather {p.get('gather_check', '-')}")


if __name__ == "__main__":
    main(sys.argv[1:])
