#!/bin/bash
# session 16: reinmav-v0 final FP64 diet: tests + timing
OUT=gpurun_out/s16
mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_reinmav.py -x -q -m gpu > "$OUT/pytest_reinmav.log" 2>&1
echo "pytest rc=$?" >> "$OUT/pytest_reinmav.log"
for rep in 1 2; do timeout 300 python tools/probe_reinmav.py >> "$OUT/probe_reinmav.log" 2>&1; done
tail -n 5 "$OUT/pytest_reinmav.log"; cat "$OUT/probe_reinmav.log"
