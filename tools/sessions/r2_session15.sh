#!/bin/bash
# session 15: reinmav-v0 regrouped; 71 registers (7 CTAs/SM) vs 64 (8 CTAs/SM, build/librmv_occ8.so)
OUT=gpurun_out/s15
mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_reinmav.py -x -q -m gpu > "$OUT/pytest_reinmav.log" 2>&1
echo "pytest rc=$?" >> "$OUT/pytest_reinmav.log"
for rep in 1 2; do
  timeout 300 python tools/probe_reinmav.py >> "$OUT/probe_reinmav.log" 2>&1
  RMV_LIB=$PWD/build/librmv_occ8.so timeout 300 python tools/probe_reinmav.py >> "$OUT/probe_reinmav_occ8.log" 2>&1
done
RMV_LIB=$PWD/build/librmv_occ8.so timeout 900 python -m pytest tests/test_reinmav.py -x -q -m gpu > "$OUT/pytest_reinmav_occ8.log" 2>&1
tail -n 5 "$OUT/pytest_reinmav.log" "$OUT/pytest_reinmav_occ8.log"
cat "$OUT/probe_reinmav.log"; echo occ8; cat "$OUT/probe_reinmav_occ8.log"
