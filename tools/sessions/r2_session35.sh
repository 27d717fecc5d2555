#!/bin/bash
# session 35: final validation of the round — sanitizers, full GPU suite, smoke, bench (default and driver flags, both arms)
OUT=gpurun_out/s35
mkdir -p "$OUT"
bash tools/sessions/r2_session27.sh > "$OUT/sanitizers.log" 2>&1
grep -E "SUMMARY|exit code" "$OUT/sanitizers.log"
timeout 1500 python -m pytest tests -x -q -m gpu > "$OUT/pytest_gpu.log" 2>&1
echo "pytest rc=$?" >> "$OUT/pytest_gpu.log"
tail -n 3 "$OUT/pytest_gpu.log"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 1
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 3 > "$OUT/bench_ref_flags.json" 2> "$OUT/bench_ref_flags.err"; echo "ref rc=$?"
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 > "$OUT/bench_flags.json" 2> "$OUT/bench_flags.err"; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("bench_ref_flags", "bench_flags"):
    d = json.loads(open("gpurun_out/s35/%s.json" % f).read().strip().splitlines()[-1])
    print(f, {k: d.get(k) for k in ("impl", "value", "steps", "warmup", "ms_per_step", "gpu_launches")}, d["e2e"]["value"], (d.get("roofline") or {}).get("frac"))
PY
