#!/bin/bash
# round 2: ncu --set full captures of every kernel family.  CSV / text pages land in gpurun_out/<dir>
# (tools/summarize_ncu.py turns them into profiles/r2_*); the bulky .ncu-rep files stay in /tmp on the GPU box.
OUT=gpurun_out/${1:-prof_r2}
REP=/tmp/rmv_ncu_reports
mkdir -p "$OUT" "$REP"
NCU="ncu --set full --cache-control none --clock-control none --import-source on"
cap() {   # cap <name> <kernel regex> <launches to skip> <command...>
  local name=$1 rx=$2 skip=$3; shift 3
  $NCU -k "regex:$rx" -s "$skip" -c 1 -f -o "$REP/$name" "$@" > "$OUT/$name.log" 2>&1
  ncu -i "$REP/$name.ncu-rep" --page raw --csv > "$OUT/$name.raw.csv" 2>/dev/null
  ncu -i "$REP/$name.ncu-rep" --page details > "$OUT/$name.details.txt" 2>/dev/null
  ncu -i "$REP/$name.ncu-rep" --page source --csv --print-source sass > "$OUT/$name.sass.csv" 2>/dev/null
}
SW="python tools/sweep.py --vec 1 --layout aos"
# the step stream (one launch = 16 / 64 steps), L2-resident and at N = 65536; forced at N = 2^23 (where the library does not use it)
cap stream_q3d_n20 k_step_stream 2 $SW --envs quadrotor3d-v0 --n 1048576 --fused 0 --steps 16 --ring 16
cap stream_q3d_n16 k_step_stream 2 $SW --envs quadrotor3d-v0 --n 65536 --fused 0 --steps 64 --ring 64
cap stream_q3dslung_n20 k_step_stream 2 $SW --envs quadrotor3d-slungload-v0 --n 1048576 --fused 0 --steps 16 --ring 16
cap stream_q3dslung_n16 k_step_stream 2 $SW --envs quadrotor3d-slungload-v0 --n 65536 --fused 0 --steps 64 --ring 64
cap stream_q2dslung_n16 k_step_stream 2 $SW --envs quadrotor2d-slungload-v0 --n 65536 --fused 0 --steps 64 --ring 64
RMV_FORCE_STREAM=1 cap stream_q3d_n23_forced k_step_stream 1 $SW --envs quadrotor3d-v0 --n 8388608 --fused 0 --steps 8 --ring 8
cap stream_pair_q3d_n20 k_step_stream_pair 2 python tools/sweep.py --vec 2 --layout aos --envs quadrotor3d-v0 --n 1048576 --fused 0 --steps 16 --ring 16
# one launch per step: the HBM-bound point and the slung-load kinds
cap step_q3d_n23 "^k_step$" 12 $SW --envs quadrotor3d-v0 --n 8388608 --fused 2 --steps 8 --ring 8
cap step_q3dslung_n20 "^k_step$" 30 $SW --envs quadrotor3d-slungload-v0 --n 1048576 --fused 2 --steps 16 --ring 16
cap step_q3dslung_n23 "^k_step$" 12 $SW --envs quadrotor3d-slungload-v0 --n 8388608 --fused 2 --steps 8 --ring 8
cap step_q2dslung_n20 "^k_step$" 30 $SW --envs quadrotor2d-slungload-v0 --n 1048576 --fused 2 --steps 16 --ring 16
cap step_q2dslung_n23 "^k_step$" 12 $SW --envs quadrotor2d-slungload-v0 --n 8388608 --fused 2 --steps 8 --ring 8
# the other kernels
cap control_q3d_n20 k_control 3 python tools/probe_misc.py control 1048576 quadrotor3d-v0
cap control_q2d_n20 k_control 3 python tools/probe_misc.py control 1048576 quadrotor2d-v0
cap control_step_q3d_n20 "^k_step$" 3 python tools/probe_misc.py control_step 1048576 quadrotor3d-v0
cap rollout_q3d_n20 k_rollout 2 python tools/probe_misc.py rollout 1048576 quadrotor3d-v0
cap rows_track_q3d_n20 k_step_rows 8 python tools/probe_misc.py rows_track 1048576 quadrotor3d-v0
cap reinmav_n20 k_reinmav_step 3 python tools/probe_misc.py reinmav 1048576
# launch list of a short bench run (serialised, cold-cache: compare shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches_bench.csv" \
    python bench.py --steps 20 --warmup 5 --no-cpu --no-extras --e2e-steps 5 --repeats 2 --min-region-ms 1 > "$OUT/launches_bench.stdout" 2> "$OUT/launches_bench.stderr"
# sanitizers
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 1 python tools/sanitize_probe.py > "$OUT/sanitizer_$tool.txt" 2>&1
  echo "exit code $?" >> "$OUT/sanitizer_$tool.txt"
done
ls -la "$OUT"
