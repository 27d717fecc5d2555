#!/bin/bash
OUT=gpurun_out/r2s3
mkdir -p $OUT
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $OUT/pytest.log 2>&1
echo "rc=$?" >> $OUT/pytest.log
timeout 300 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1
echo "smoke rc=$?" >> $OUT/smoke.log
S="timeout 600 python tools/sweep.py --layout aos --vec 1"
$S --envs quadrotor3d-v0 --n 65536,262144,524288,1048576,2097152,4194304,8388608 --fused 0,2 --steps 256 --ring 32 > $OUT/sweep_q3d.txt 2>&1
$S --envs quadrotor3d-v0 --n 65536 --fused 0 --sblock 32,64,128 --steps 2048 --ring 64 > $OUT/sweep_q3d_n16_blocks.txt 2>&1
$S --envs quadrotor3d-v0 --n 262144 --fused 0 --sblock 32,64,128 --steps 1024 --ring 64 > $OUT/sweep_q3d_n18_blocks.txt 2>&1
$S --envs quadrotor2d-v0,quadrotor2d-slungload-v0,quadrotor3d-slungload-v0 --n 65536,1048576,8388608 --fused 0,2 --steps 128 --ring 16 > $OUT/sweep_other_kinds.txt 2>&1
$S --envs quadrotor3d-v0 --n 1048576,65536 --fused 0,2 --track 1 --steps 256 --ring 32 > $OUT/sweep_q3d_track.txt 2>&1
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err
echo "bench rc=$?" >> $OUT/bench.err
ncu --set full --cache-control none --clock-control none --import-source on -k regex:k_step_stream -s 2 -c 1 -f -o $OUT/stream_q3d_n20 \
    python tools/sweep.py --envs quadrotor3d-v0 --n 1048576 --vec 1 --layout aos --fused 0 --steps 16 --ring 16 > $OUT/ncu_a.log 2>&1
ncu -i $OUT/stream_q3d_n20.ncu-rep --page raw --csv > $OUT/stream_q3d_n20.raw.csv 2>/dev/null
ncu -i $OUT/stream_q3d_n20.ncu-rep --page details > $OUT/stream_q3d_n20.details.txt 2>/dev/null
ncu -i $OUT/stream_q3d_n20.ncu-rep --page source --csv --print-source sass > $OUT/stream_q3d_n20.sass.csv 2>/dev/null
rm -f $OUT/stream_q3d_n20.ncu-rep
