#!/bin/bash
OUT=gpurun_out/r2s2
mkdir -p $OUT
( time timeout 1200 python -m pytest tests/test_gpu_vec_env.py tests/test_launcher_seam.py tests/test_capi_symbols.py -m gpu -x -q ) > $OUT/pytest_vec.log 2>&1
echo "rc=$?" >> $OUT/pytest_vec.log
( time timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stream or fused or graph_replay or full_size" ) > $OUT/pytest_stream.log 2>&1
echo "rc=$?" >> $OUT/pytest_stream.log
timeout 300 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1
echo "smoke rc=$?" >> $OUT/smoke.log
S="timeout 600 python tools/sweep.py --layout aos --vec 1"
$S --envs quadrotor3d-v0 --n 65536,262144,524288,1048576,2097152,8388608 --fused 0,2 --steps 256 --ring 32 > $OUT/sweep_q3d.txt 2>&1
$S --envs quadrotor3d-v0 --n 1048576 --fused 0 --warps 0,56,48 --sblock 128,64 --steps 256 --ring 64 > $OUT/sweep_q3d_n20_shapes.txt 2>&1
$S --envs quadrotor2d-v0,quadrotor2d-slungload-v0,quadrotor3d-slungload-v0 --n 65536,1048576,8388608 --fused 0,2 --steps 128 --ring 16 > $OUT/sweep_other_kinds.txt 2>&1
$S --envs quadrotor3d-v0 --n 1048576,65536 --fused 0,2 --track 1 --steps 256 --ring 32 > $OUT/sweep_q3d_track.txt 2>&1
ncu --set full --cache-control none --clock-control none --import-source on -k regex:k_step_stream -s 2 -c 1 -f -o $OUT/stream_q3d_n20 \
    python tools/sweep.py --envs quadrotor3d-v0 --n 1048576 --vec 1 --layout aos --fused 0 --steps 16 --ring 16 > $OUT/ncu_a.log 2>&1
ncu -i $OUT/stream_q3d_n20.ncu-rep --page raw --csv > $OUT/stream_q3d_n20.raw.csv 2>/dev/null
ncu -i $OUT/stream_q3d_n20.ncu-rep --page details > $OUT/stream_q3d_n20.details.txt 2>/dev/null
ncu -i $OUT/stream_q3d_n20.ncu-rep --page source --csv --print-source sass > $OUT/stream_q3d_n20.sass.csv 2>/dev/null
rm -f $OUT/stream_q3d_n20.ncu-rep
