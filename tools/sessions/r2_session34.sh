#!/bin/bash
# session 34: ncu --set full of the step stream on a row-state handle
OUT=gpurun_out/prof_r2d
REP=/tmp/rmv_ncu_reports
mkdir -p "$OUT" "$REP"
NCU="ncu --set full --cache-control none --clock-control none --import-source on"
cap() { local name=$1 rx=$2 skip=$3; shift 3
  $NCU -k "regex:$rx" -s "$skip" -c 1 -f -o "$REP/$name" "$@" > "$OUT/$name.log" 2>&1
  ncu -i "$REP/$name.ncu-rep" --page raw --csv > "$OUT/$name.raw.csv" 2>/dev/null
  ncu -i "$REP/$name.ncu-rep" --page details > "$OUT/$name.details.txt" 2>/dev/null
  ncu -i "$REP/$name.ncu-rep" --page source --csv --print-source sass > "$OUT/$name.sass.csv" 2>/dev/null
  grep -E "Duration|Issue Slots Busy|Registers Per|Achieved Occupancy|Shared Memory Configuration" "$OUT/$name.details.txt" | head -6
}
SW="python tools/sweep.py --vec 1 --layout aos --rows 3"
cap stream_rows_q3d_n16 k_step_stream 2 $SW --envs quadrotor3d-v0 --n 65536 --fused 0 --steps 64 --ring 64
cap stream_rows_q3d_n18 k_step_stream 2 $SW --envs quadrotor3d-v0 --n 262144 --fused 0 --steps 32 --ring 32
cap stream_rows_q3dslung_n18 k_step_stream 2 $SW --envs quadrotor3d-slungload-v0 --n 262144 --fused 0 --steps 32 --ring 32
