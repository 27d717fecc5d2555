#!/bin/bash
OUT=gpurun_out/s26
mkdir -p "$OUT"
for v in 0 1; do
  RMV_STREAM_PAIR=$v timeout 600 python bench.py --no-cpu --no-extras > "$OUT/bench_pair$v.json" 2> "$OUT/bench_pair$v.err"
done
python - <<'PY'
import json
for v in (0, 1):
    d = json.loads(open("gpurun_out/s26/bench_pair%d.json" % v).read().strip().splitlines()[-1])
    print("pair", v, d["value"], d["ms_per_step"], d["clocks"], d["config"]["timing"][-60:])
PY
nvidia-smi --query-gpu=power.limit,power.max_limit,clocks.max.sm,clocks.max.mem --format=csv
