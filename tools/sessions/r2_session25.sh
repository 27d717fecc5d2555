#!/bin/bash
# session 25: full GPU suite, smoke, ncu of the norm kernels as committed, bench
OUT=gpurun_out/s25
PR=gpurun_out/prof_r2c
REP=/tmp/rmv_ncu_reports
mkdir -p "$OUT" "$PR" "$REP"
timeout 1500 python -m pytest tests -x -q -m gpu > "$OUT/pytest_gpu.log" 2>&1
echo "pytest rc=$?" >> "$OUT/pytest_gpu.log"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > "$OUT/smoke.log" 2>&1
NCU="ncu --set full --cache-control none --clock-control none --import-source on"
for k in k_norm_moments k_norm_apply; do
  $NCU -k "regex:$k" -s 20 -c 1 -f -o "$REP/${k}_q3d_n20" python tools/probe_misc.py vecnorm 1048576 quadrotor3d-v0 > "$PR/${k}_q3d_n20.log" 2>&1
  ncu -i "$REP/${k}_q3d_n20.ncu-rep" --page raw --csv > "$PR/${k}_q3d_n20.raw.csv" 2>/dev/null
  ncu -i "$REP/${k}_q3d_n20.ncu-rep" --page details > "$PR/${k}_q3d_n20.details.txt" 2>/dev/null
  ncu -i "$REP/${k}_q3d_n20.ncu-rep" --page source --csv --print-source sass > "$PR/${k}_q3d_n20.sass.csv" 2>/dev/null
done
timeout 900 python bench.py > "$OUT/bench.json" 2> "$OUT/bench.err"
tail -n 4 "$OUT/pytest_gpu.log"; cat "$OUT/smoke.log" | tail -n 2
python - <<'PY'
import json
d = json.loads(open("gpurun_out/s25/bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"]["value"], d["roofline"]["frac"], d["cpu_baseline"])
for k in ("hbm_bound", "cold_l2", "n65536_quadrotor3d", "n65536_quadrotor3d_slungload"):
    print(k, d["roofline"][k]["us_per_step"], d["roofline"][k]["frac"])
for k, v in d["extras"].items():
    print(k, round(v["us_per_step"], 2))
PY
