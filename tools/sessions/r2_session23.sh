#!/bin/bash
OUT=gpurun_out/prof_r2c
REP=/tmp/rmv_ncu_reports
mkdir -p "$OUT" "$REP"
NCU="ncu --set full --cache-control none --clock-control none --import-source on"
for k in k_norm_moments k_norm_apply; do
  $NCU -k "regex:$k" -s 20 -c 1 -f -o "$REP/${k}_q3d_n20" python tools/probe_misc.py vecnorm 1048576 quadrotor3d-v0 > "$OUT/${k}_q3d_n20.log" 2>&1
  ncu -i "$REP/${k}_q3d_n20.ncu-rep" --page raw --csv > "$OUT/${k}_q3d_n20.raw.csv" 2>/dev/null
  ncu -i "$REP/${k}_q3d_n20.ncu-rep" --page details > "$OUT/${k}_q3d_n20.details.txt" 2>/dev/null
  ncu -i "$REP/${k}_q3d_n20.ncu-rep" --page source --csv --print-source sass > "$OUT/${k}_q3d_n20.sass.csv" 2>/dev/null
  grep -E "Duration|Issue Slots Busy|DRAM Throughput|Memory Throughput|FP64|Registers Per|Achieved Occupancy|No Eligible|L2 Hit" "$OUT/${k}_q3d_n20.details.txt" | head -12
done
timeout 300 python tools/probe_misc.py vecnorm 1048576 quadrotor3d-v0
