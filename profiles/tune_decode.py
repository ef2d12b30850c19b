"""Kernel-tuning sweep: time decode_compact (LDG variant) for every alternative build under build/variants/.
Each build is measured in its own process (YOLO_B200_LIB selects the library)."""
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SNIPPET = r'''
import sys, torch
sys.path.insert(0, %r)
from pytorch_yolo_b200 import ops, synth
wl, B, conf, variant = %r, %d, %f, %r
w = synth.WORKLOADS[wl]
heads = synth.synth_heads(wl, B, "B", seed=1234, device="cuda:0")
specs = [ops.scale_spec(a, g, g, w["img_size"]) for a, g in zip(w["anchors"], w["grids"])]
buf = ops.Buffers("cuda:0", B, synth.anchors_per_image(wl), w["nc"])
for _ in range(5): ops.decode_compact(heads, specs, w["nc"], conf, buf, variant=variant)
torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): ops.decode_compact(heads, specs, w["nc"], conf, buf, variant=variant)
    e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 50)
nbytes = B * synth.head_bytes_per_image(wl)
print(f"{best*1000:8.1f} us {nbytes/best/1e6:8.1f} GB/s")
'''


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "spp-608"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    conf = float(sys.argv[3]) if len(sys.argv) > 3 else 0.3
    libs = sorted(glob.glob(os.path.join(ROOT, "build", "variants", "*.so")))
    for lib in [None] + libs:
        env = dict(os.environ)
        if lib:
            env["YOLO_B200_LIB"] = lib
        for variant in (("ldg", "tma", "tma2d") if lib is None else ("ldg",)):
            r = subprocess.run([sys.executable, "-c", SNIPPET % (ROOT, wl, B, conf, variant)], env=env,
                               capture_output=True, text=True, timeout=300)
            name = os.path.basename(lib) if lib else "default"
            print(f"{name:24s} {variant:4s} {r.stdout.strip() or r.stderr.strip()[-200:]}", flush=True)


if __name__ == "__main__":
    main()
