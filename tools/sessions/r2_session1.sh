#!/bin/bash
# round-2 GPU session 1: parity first, then the step-stream sweep and a bench line
OUT=gpurun_out/r2s1
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $OUT/pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest.log
timeout 300 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1
echo "smoke rc=$?" >> $OUT/smoke.log
S="timeout 600 python tools/sweep.py --layout aos --vec 1"
$S --envs quadrotor3d-v0 --n 1048576 --fused 0 --pipe 0,1 --warps 0,48,40,32,24,16 --steps 256 --ring 64 > $OUT/sweep_q3d_n20_stream.txt 2>&1
$S --envs quadrotor3d-v0 --n 1048576 --fused 2,1 --steps 256 --ring 64 > $OUT/sweep_q3d_n20_ref.txt 2>&1
$S --envs quadrotor3d-v0 --n 65536 --fused 0 --pipe 0,1 --sblock 32,64,128 --steps 4096 --ring 64 > $OUT/sweep_q3d_n16_stream.txt 2>&1
$S --envs quadrotor3d-v0 --n 65536 --fused 2,1 --steps 4096 --ring 64 > $OUT/sweep_q3d_n16_ref.txt 2>&1
$S --envs quadrotor3d-v0 --n 8388608 --fused 0,2 --pipe 0,1 --steps 64 --ring 8 > $OUT/sweep_q3d_n23.txt 2>&1
$S --envs quadrotor3d-v0 --n 262144,524288,2097152 --fused 0,2 --pipe 0,1 --steps 256 --ring 32 > $OUT/sweep_q3d_mid.txt 2>&1
$S --envs quadrotor2d-v0,quadrotor2d-slungload-v0,quadrotor3d-slungload-v0 --n 1048576,65536 --fused 0,2 --pipe 0,1 --steps 256 --ring 32 > $OUT/sweep_other_kinds.txt 2>&1
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err
echo "bench rc=$?" >> $OUT/bench.err
