#!/bin/bash
# One bench line per BASELINE.json config (1 GPU): writes gpurun_out/all_configs.jsonl
set -e
cd "$(dirname "$0")/.."
out=gpurun_out/all_configs.jsonl
mkdir -p gpurun_out; : > $out
common="--steps 400 --warmup 40 --no-second-flow"
python bench.py $common --mode SN --batch 4 --no-train-step >> $out 2>/dev/null                  # configs[0] (its GPU timing)
python bench.py $common >> $out 2>/dev/null                                                      # configs[1] + configs[2] (train_step)
python bench.py $common --mode TG --no-cpu-baseline --no-train-step >> $out 2>/dev/null          # configs[2]'s loss path (TG)
python bench.py $common --mode DS --no-cpu-baseline --no-train-step >> $out 2>/dev/null          # configs[3]
python bench.py $common --mode DC --no-cpu-baseline --no-train-step >> $out 2>/dev/null          # configs[3]
python bench.py --steps 200 --warmup 20 --no-second-flow --shape 375x1242 --no-train-step >> $out 2>/dev/null   # configs[4]
python - <<'PY'
import json
for line in open("gpurun_out/all_configs.jsonl"):
    d = json.loads(line)
    c = d["config"]
    print("%-4s B=%-2d %dx%d scales=%d : %8.0f frames/s  %7.1f us/step  e2e %7.0f frames/s  frac %.3f  cpu %s  train %s" % (
        c["mode"], c["batch_per_gpu"], c["height"], c["width"], len(c["scales"]), d["value"], d["ms_per_step"] * 1e3, d["e2e"]["value"],
        d["roofline"]["frac"], ("%.1f" % d["cpu_baseline"]["value"]) if d.get("cpu_baseline") else "-",
        ("%.0f" % d["train_step"]["value"]) if d.get("train_step") and "value" in d["train_step"] else "-"))
PY
