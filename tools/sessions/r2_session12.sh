#!/bin/bash
# session 12: FULL-unit fast path of the step stream — parity subset + sweep against the committed numbers
OUT=gpurun_out/s12
mkdir -p "$OUT"
timeout 900 python -m pytest tests -x -q -m gpu -k "stream or full_size or seam" > "$OUT/pytest_stream.log" 2>&1
echo "pytest rc=$?" >> "$OUT/pytest_stream.log"
for rep in 1 2 3; do
  for env in quadrotor3d-v0 quadrotor3d-slungload-v0 quadrotor2d-v0 quadrotor2d-slungload-v0; do
    timeout 300 python tools/sweep.py --vec 1 --layout aos --envs $env --n 65536,1048576 --fused 0 >> "$OUT/sweep_stream.log" 2>&1
  done
done
timeout 600 python bench.py --no-cpu > "$OUT/bench.json" 2> "$OUT/bench.err"
tail -n 40 "$OUT/sweep_stream.log"
tail -n 3 "$OUT/pytest_stream.log"
cat "$OUT/bench.json"
