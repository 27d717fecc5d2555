#!/bin/bash
# session 27: sanitizers on the extended probe
OUT=gpurun_out/s27
mkdir -p "$OUT"
timeout 300 python tools/sanitize_probe.py > "$OUT/probe_plain.txt" 2>&1
echo "exit code $?" >> "$OUT/probe_plain.txt"
tail -n 14 "$OUT/probe_plain.txt"
grep -q "sanitize probe finished" "$OUT/probe_plain.txt" || exit 1
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 1 python tools/sanitize_probe.py > "$OUT/sanitizer_$tool.txt" 2>&1
  echo "exit code $?" >> "$OUT/sanitizer_$tool.txt"
  tail -n 4 "$OUT/sanitizer_$tool.txt"
done
