#!/bin/bash
# Runs bench.py on N GPUs for the BASELINE.json configs; one JSON line per run into gpurun_out/configs_n$N.jsonl
#   bash profiles/run_configs.sh N [configs...]   (configs: cfg2 cfg3 cfg4 cfg5)
N=$1; shift
CFGS=${@:-cfg2 cfg3 cfg4 cfg5}
OUT=gpurun_out/configs_n$N.jsonl
: > $OUT
run() {
  if [ "$N" = "1" ]; then python bench.py --gpus 1 "$@" 2>>gpurun_out/configs_n$N.err | tail -1 >> $OUT
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29540 bench.py --gpus $N "$@" 2>>gpurun_out/configs_n$N.err | tail -1 >> $OUT; fi
}
for c in $CFGS; do
  case $c in
    cfg2) run --workload spp-608 --batch 64 --conf 0.3 --steps 1000 --warmup 10 ;;
    cfg3) run --workload tiny-416 --batch 1024 --conf 0.3 --steps 500 --warmup 10 --no-e2e --no-cpu-baseline ;;
    cfg4) run --workload spp-608 --batch 64 --conf 0.001 --steps 300 --warmup 10 --no-e2e --no-cpu-baseline ;;
    cfg5) run --workload spp-1024 --batch 256 --conf 0.3 --steps 200 --warmup 5 --no-e2e --no-cpu-baseline ;;
  esac
done
python - <<PY
import json
for l in open("$OUT"):
    try: d = json.loads(l)
    except Exception: print("bad line:", l[:200]); continue
    print(d["n_gpus"], d["config"]["workload"][:60], "| img/s", round(d["value"]), "| ms/step", round(d["ms_per_step"], 4),
          "| roofline", round(d["roofline"]["frac"], 3), "| e2e", round(d.get("e2e", {}).get("value", 0)))
PY
