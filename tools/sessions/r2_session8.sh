#!/bin/bash
OUT=gpurun_out/r2s8
mkdir -p $OUT
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stream or full_size" ) > $OUT/pytest_stream.log 2>&1
echo "rc=$?" >> $OUT/pytest_stream.log
S="timeout 900 python tools/sweep.py --layout aos --vec 1"
$S --envs quadrotor3d-v0 --n 65536,262144,1048576,2097152 --fused 0,2 --steps 256 --ring 32 > $OUT/sweep_q3d.txt 2>&1
$S --envs quadrotor2d-v0,quadrotor2d-slungload-v0,quadrotor3d-slungload-v0 --n 65536,262144,1048576 --fused 0 --steps 256 --ring 32 > $OUT/sweep_other.txt 2>&1
$S --envs quadrotor3d-v0 --n 1048576,65536 --fused 0 --track 1 --steps 256 --ring 32 > $OUT/sweep_q3d_track.txt 2>&1
