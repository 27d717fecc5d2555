#!/bin/bash
OUT=gpurun_out/r2s10
mkdir -p $OUT
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stream or full_size" ) > $OUT/pytest_stream.log 2>&1
echo "rc=$?" >> $OUT/pytest_stream.log
S="timeout 900 python tools/sweep.py --layout aos --vec 1"
$S --envs quadrotor3d-v0 --n 65536,131072,262144,524288,1048576 --fused 0 --apre 0,1 --steps 512 --ring 64 > $OUT/sweep_q3d_apre.txt 2>&1
$S --envs quadrotor2d-v0,quadrotor2d-slungload-v0,quadrotor3d-slungload-v0 --n 65536,262144,1048576 --fused 0 --apre 0,1 --steps 256 --ring 32 > $OUT/sweep_other_apre.txt 2>&1
$S --envs quadrotor3d-v0 --n 65536 --fused 0 --apre 1 --sblock 32,64,128 --steps 2048 --ring 64 > $OUT/sweep_q3d_n16_apre_blocks.txt 2>&1
