#!/bin/bash
# session 33: final validation — full GPU suite, smoke, the default bench line
OUT=gpurun_out/s33
mkdir -p "$OUT"
timeout 1500 python -m pytest tests -x -q -m gpu > "$OUT/pytest_gpu.log" 2>&1
echo "pytest rc=$?" >> "$OUT/pytest_gpu.log"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > "$OUT/smoke.log" 2>&1
timeout 900 python bench.py > "$OUT/bench.json" 2> "$OUT/bench.err"
echo "bench rc=$?"
tail -n 4 "$OUT/pytest_gpu.log"; tail -n 2 "$OUT/smoke.log"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/s33/bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "clocks")}, d["e2e"]["value"], d["roofline"]["frac"], d["cpu_baseline"]["value"])
for k in ("hbm_bound", "cold_l2", "n65536_quadrotor3d", "n65536_quadrotor3d_slungload"):
    print(k, d["roofline"][k]["us_per_step"], d["roofline"][k]["frac"])
for k, v in d["extras"].items():
    print(k, round(v["us_per_step"], 2))
PY
