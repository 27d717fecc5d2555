#!/bin/bash
OUT=gpurun_out/s20
mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_reinmav.py -x -q -m gpu > "$OUT/pytest_reinmav.log" 2>&1
echo "pytest rc=$?" >> "$OUT/pytest_reinmav.log"
for rep in 1 2; do
  timeout 300 python tools/probe_reinmav.py >> "$OUT/probe_reinmav.log" 2>&1
  RMV_LIB=$PWD/build/librmv_r78.so timeout 300 python tools/probe_reinmav.py >> "$OUT/probe_reinmav_r78.log" 2>&1
done
tail -n 5 "$OUT/pytest_reinmav.log"
cat "$OUT/probe_reinmav.log"; echo r78; cat "$OUT/probe_reinmav_r78.log"
