#!/bin/bash
# Captures the evidence under profiles/ for the current build (run on the GPU box through gpurun).
# usage: tools/profile_round.sh <out-dir under gpurun_out/>
set -x
OUT=gpurun_out/${1:-prof}
mkdir -p "$OUT"
NCU="ncu --clock-control none"
# 1. launch list of a short bench run (serialised, cold-cache: compare shares, not absolutes)
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file "$OUT/launches_bench.csv" \
    python bench.py --steps 40 --warmup 3 --no-cpu --no-extras --e2e-steps 5 --repeats 1 > "$OUT/launches_bench.stdout" 2> "$OUT/launches_bench.stderr"
# 2. the headline kernel, cold (default cache control) and warm (steady state of a back-to-back run)
$NCU --set full --import-source on -k regex:k_step -s 12 -c 1 -f -o "$OUT/q3d_n20_cold" \
    python tools/sweep.py --envs quadrotor3d-v0 --vec 1 --block 0 --layout aos --steps 8 --ring 4 > "$OUT/a.log" 2>&1
$NCU --set full --cache-control none --import-source on -k regex:k_step -s 30 -c 1 -f -o "$OUT/q3d_n20_warm" \
    python tools/sweep.py --envs quadrotor3d-v0 --vec 1 --block 0 --layout aos --steps 16 --ring 16 > "$OUT/b.log" 2>&1
$NCU --set full --cache-control none --import-source on -k regex:k_step -s 12 -c 1 -f -o "$OUT/q3d_n23" \
    python tools/sweep.py --envs quadrotor3d-v0 --n 8388608 --vec 1 --block 0 --layout aos --steps 8 --ring 8 > "$OUT/c.log" 2>&1
# 3. the planar kind (CTA 256 + prefetch when L2-resident, 128 + prefetch when streaming)
$NCU --set full --cache-control none --import-source on -k regex:k_step -s 30 -c 1 -f -o "$OUT/q2d_n20_warm" \
    python tools/sweep.py --envs quadrotor2d-v0 --vec 1 --block 0 --layout aos --steps 16 --ring 16 > "$OUT/d.log" 2>&1
$NCU --set full --cache-control none --import-source on -k regex:k_step -s 12 -c 1 -f -o "$OUT/q2d_n23" \
    python tools/sweep.py --envs quadrotor2d-v0 --n 8388608 --vec 1 --block 0 --layout aos --steps 8 --ring 8 > "$OUT/e.log" 2>&1
# ncu reports are large (gpurun brings back <= 64 MiB): keep the CSV pages, drop all reports but the headline one
for r in q3d_n20_cold q3d_n20_warm q3d_n23 q2d_n20_warm q2d_n23; do
  ncu -i "$OUT/$r.ncu-rep" --page raw --csv > "$OUT/$r.raw.csv" 2>/dev/null
  ncu -i "$OUT/$r.ncu-rep" --page source --csv --print-source sass > "$OUT/$r.sass.csv" 2>/dev/null
  ncu -i "$OUT/$r.ncu-rep" --page details > "$OUT/$r.details.txt" 2>/dev/null
  [ "$r" = q3d_n20_warm ] || rm -f "$OUT/$r.ncu-rep"
done
# 4. sanitizers on a small workload that touches every kernel variant
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 1 python tools/sanitize_probe.py > "$OUT/sanitizer_$tool.txt" 2>&1
  echo "exit code $?" >> "$OUT/sanitizer_$tool.txt"
done
ls -la "$OUT"
