#!/bin/bash
OUT=gpurun_out/s21
mkdir -p "$OUT"
build/f64_check > "$OUT/f64_check.txt" 2>&1
timeout 900 python -m pytest tests/test_reinmav.py -x -q -m gpu > "$OUT/pytest_reinmav.log" 2>&1
echo "pytest rc=$?" >> "$OUT/pytest_reinmav.log"
for rep in 1 2; do timeout 300 python tools/probe_reinmav.py >> "$OUT/probe_reinmav.log" 2>&1; done
cat "$OUT/f64_check.txt"; tail -n 5 "$OUT/pytest_reinmav.log"; cat "$OUT/probe_reinmav.log"
