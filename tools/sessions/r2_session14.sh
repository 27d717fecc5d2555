#!/bin/bash
# session 14: ncu --set full of k_reinmav_step after the FP64 diet
OUT=gpurun_out/s14
REP=/tmp/rmv_ncu_reports
mkdir -p "$OUT" "$REP"
NCU="ncu --set full --cache-control none --clock-control none --import-source on"
$NCU -k "regex:k_reinmav_step" -s 3 -c 1 -f -o "$REP/reinmav_n20" python tools/probe_misc.py reinmav 1048576 > "$OUT/reinmav_n20.log" 2>&1
ncu -i "$REP/reinmav_n20.ncu-rep" --page raw --csv > "$OUT/reinmav_n20.raw.csv" 2>/dev/null
ncu -i "$REP/reinmav_n20.ncu-rep" --page details > "$OUT/reinmav_n20.details.txt" 2>/dev/null
ncu -i "$REP/reinmav_n20.ncu-rep" --page source --csv --print-source sass > "$OUT/reinmav_n20.sass.csv" 2>/dev/null
grep -E "Duration|Issue Slots Busy|Executed Ipc|FP64|Registers Per|Achieved Occupancy|Theoretical Occ|Stall|No Eligible|Eligible Warps|Issued Warp" "$OUT/reinmav_n20.details.txt" | head -40
