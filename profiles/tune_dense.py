"""Sweep alternative builds (build/variants/*.so) timing decode_dense only."""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SNIPPET = r'''
import sys, torch
sys.path.insert(0, %r)
from pytorch_yolo_b200 import ops, synth
wl, B = %r, %d
w = synth.WORKLOADS[wl]
heads = synth.synth_heads(wl, B, "B", seed=1234, device="cuda:0")
specs = [ops.scale_spec(a, g, g, w["img_size"]) for a, g in zip(w["anchors"], w["grids"])]
pred = torch.empty(B, synth.anchors_per_image(wl), w["nc"] + 5, device="cuda:0")
for _ in range(5): ops.decode_dense(heads, specs, w["nc"], out=pred)
torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30): ops.decode_dense(heads, specs, w["nc"], out=pred)
    e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 30)
nbytes = 2 * B * synth.head_bytes_per_image(wl)
print(f"{best*1000:8.1f} us {nbytes/best/1e6:8.1f} GB/s")
'''
wl = sys.argv[1] if len(sys.argv) > 1 else "spp-608"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
for lib in sorted(glob.glob(os.path.join(ROOT, "build", "variants", "dd_*.so"))):
    env = dict(os.environ, YOLO_B200_LIB=lib)
    r = subprocess.run([sys.executable, "-c", SNIPPET % (ROOT, wl, B)], env=env, capture_output=True, text=True, timeout=300)
    print(f"{os.path.basename(lib):20s} {r.stdout.strip() or r.stderr.strip()[-200:]}", flush=True)
