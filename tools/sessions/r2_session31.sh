#!/bin/bash
# session 31: full GPU suite + smoke on the final code
OUT=gpurun_out/s31
mkdir -p "$OUT"
timeout 1500 python -m pytest tests -x -q -m gpu > "$OUT/pytest_gpu.log" 2>&1
echo "pytest rc=$?" >> "$OUT/pytest_gpu.log"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > "$OUT/smoke.log" 2>&1
tail -n 5 "$OUT/pytest_gpu.log"; tail -n 2 "$OUT/smoke.log"
