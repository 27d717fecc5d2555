#!/bin/bash
OUT=gpurun_out/s29
mkdir -p "$OUT"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > "$OUT/bench_n2.json" 2> "$OUT/bench_n2.err"
echo "rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 20 --warmup 3 > "$OUT/bench_ref_n2.json" 2> "$OUT/bench_ref_n2.err"
echo "rc=$?"
python - <<'PY'
import json
for f in ("bench_n2", "bench_ref_n2"):
    lines = [l for l in open("gpurun_out/s29/%s.json" % f).read().strip().splitlines() if l.startswith("{")]
    print(f, len(lines))
    d = json.loads(lines[-1])
    print({k: d.get(k) for k in ("impl", "value", "n_gpus", "ms_per_step", "clocks", "gpu_launches")}, d["e2e"]["value"], d.get("episode_stats"))
PY
tail -n 3 "$OUT/bench_n2.err"
