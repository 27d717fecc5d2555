#!/bin/bash
OUT=gpurun_out/s24
mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_vecnorm.py -x -q -m gpu > "$OUT/pytest_vecnorm.log" 2>&1
echo "pytest rc=$?" >> "$OUT/pytest_vecnorm.log"
for env in quadrotor3d-v0 quadrotor2d-v0 quadrotor3d-slungload-v0; do timeout 300 python tools/probe_misc.py vecnorm 1048576 $env >> "$OUT/vecnorm.log" 2>&1; done
timeout 300 python tools/probe_misc.py vecnorm 65536 quadrotor3d-v0 >> "$OUT/vecnorm.log" 2>&1
tail -n 8 "$OUT/pytest_vecnorm.log"; cat "$OUT/vecnorm.log"
bash tools/sessions/r2_session23.sh 2>&1 | grep -E "Duration|Issue Slots|DRAM Throughput"
