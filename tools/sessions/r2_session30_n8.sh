#!/bin/bash
# session 30: both arms exactly as the driver launches them at N = 8
OUT=gpurun_out/s30
mkdir -p "$OUT"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 3 > "$OUT/bench_n8.json" 2> "$OUT/bench_n8.err"
echo "rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus 8 --steps 20 --warmup 3 > "$OUT/bench_ref_n8.json" 2> "$OUT/bench_ref_n8.err"
echo "rc=$?"
python - <<'PY'
import json
for f in ("bench_n8", "bench_ref_n8"):
    lines = [l for l in open("gpurun_out/s30/%s.json" % f).read().strip().splitlines() if l.startswith("{")]
    d = json.loads(lines[-1])
    print(f, len(lines), {k: d.get(k) for k in ("impl", "value", "n_gpus", "ms_per_step", "clocks", "gpu_launches")}, d["e2e"]["value"], d.get("episode_stats"))
PY
