#!/bin/bash
# session 13: reinmav-v0 with the short float64 asin / atan2 and hand-contracted FMAs
OUT=gpurun_out/s13
mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_reinmav.py -x -q -m gpu > "$OUT/pytest_reinmav.log" 2>&1
echo "pytest rc=$?" >> "$OUT/pytest_reinmav.log"
for rep in 1 2 3; do timeout 300 python tools/probe_reinmav.py >> "$OUT/probe_reinmav.log" 2>&1; done
tail -n 15 "$OUT/pytest_reinmav.log"
cat "$OUT/probe_reinmav.log"
