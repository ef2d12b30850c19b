"""Sweep alternative builds (build/variants/*.so): full pipelined step time via bench.py."""
import glob, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for lib in [None] + sorted(glob.glob(os.path.join(ROOT, "build", "variants", "*.so"))):
    env = dict(os.environ)
    if lib:
        env["YOLO_B200_LIB"] = lib
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--no-cpu-baseline", "--no-e2e"] + sys.argv[1:],
                       env=env, capture_output=True, text=True, timeout=600)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        print(f"{os.path.basename(lib) if lib else 'default':16s} {d['value']:10.0f} img/s  {d['ms_per_step']*1000:7.1f} us/step  "
              f"decode alone {d['roofline']['kernel_ms']*1000:6.1f} us", flush=True)
    except Exception as e:  # noqa: BLE001
        print(lib, "failed", r.stderr[-300:])
