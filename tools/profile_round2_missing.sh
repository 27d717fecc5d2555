#!/bin/bash
# the captures of tools/profile_round2.sh whose kernel filter did not match in the first pass (one launch per step kernels)
OUT=gpurun_out/${1:-prof_r2b}
REP=/tmp/rmv_ncu_reports
mkdir -p "$OUT" "$REP"
NCU="ncu --set full --cache-control none --clock-control none --import-source on"
cap() {
  local name=$1 rx=$2 skip=$3; shift 3
  $NCU -k "regex:$rx" -s "$skip" -c 1 -f -o "$REP/$name" "$@" > "$OUT/$name.log" 2>&1
  ncu -i "$REP/$name.ncu-rep" --page raw --csv > "$OUT/$name.raw.csv" 2>/dev/null
  ncu -i "$REP/$name.ncu-rep" --page details > "$OUT/$name.details.txt" 2>/dev/null
  ncu -i "$REP/$name.ncu-rep" --page source --csv --print-source sass > "$OUT/$name.sass.csv" 2>/dev/null
}
SW="python tools/sweep.py --vec 1 --layout aos"
cap step_q3d_n23 '^k_step$' 12 $SW --envs quadrotor3d-v0 --n 8388608 --fused 2 --steps 8 --ring 8
cap step_q3d_n20 '^k_step$' 30 $SW --envs quadrotor3d-v0 --n 1048576 --fused 2 --steps 16 --ring 16
cap step_q3dslung_n20 '^k_step$' 30 $SW --envs quadrotor3d-slungload-v0 --n 1048576 --fused 2 --steps 16 --ring 16
cap step_q3dslung_n23 '^k_step$' 12 $SW --envs quadrotor3d-slungload-v0 --n 8388608 --fused 2 --steps 8 --ring 8
cap step_q2dslung_n20 '^k_step$' 30 $SW --envs quadrotor2d-slungload-v0 --n 1048576 --fused 2 --steps 16 --ring 16
cap step_q2dslung_n23 '^k_step$' 12 $SW --envs quadrotor2d-slungload-v0 --n 8388608 --fused 2 --steps 8 --ring 8
cap control_step_q3d_n20 '^k_step$' 3 python tools/probe_misc.py control_step 1048576 quadrotor3d-v0
ls -la "$OUT"
