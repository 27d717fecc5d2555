#!/bin/bash
# session 32: the step stream on row-state handles — parity, then stream (0) against launches (2)
OUT=gpurun_out/s32
mkdir -p "$OUT"
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_vec_env.py -x -q -m gpu -k "row_state or stream" > "$OUT/pytest_rows_stream.log" 2>&1
echo "pytest rc=$?" >> "$OUT/pytest_rows_stream.log"
tail -n 6 "$OUT/pytest_rows_stream.log"
for env in quadrotor3d-v0 quadrotor3d-slungload-v0 quadrotor2d-v0; do
  for track in 0 1; do
    timeout 300 python tools/sweep.py --vec 1 --layout aos --envs $env --n 65536,262144,1048576 --fused 0,2 --rows 3 --track $track >> "$OUT/sweep_rows_stream.log" 2>&1
  done
done
grep -v "^env" "$OUT/sweep_rows_stream.log"
