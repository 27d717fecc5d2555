#!/bin/bash
OUT=gpurun_out/r2s4
mkdir -p $OUT
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stream" ) > $OUT/pytest_stream.log 2>&1
echo "rc=$?" >> $OUT/pytest_stream.log
( time timeout 900 python -m pytest tests/test_gpu_vec_env.py -m gpu -x -q ) > $OUT/pytest_vec.log 2>&1
echo "rc=$?" >> $OUT/pytest_vec.log
export RMV_FORCE_STREAM=1
S="timeout 900 python tools/sweep.py --layout aos"
$S --envs quadrotor3d-v0 --n 65536,262144,524288,1048576,2097152,8388608 --vec 1,2 --fused 0 --steps 256 --ring 32 > $OUT/sweep_q3d_pair.txt 2>&1
$S --envs quadrotor3d-v0 --n 1048576 --vec 2 --fused 0 --warps 0,28,24 --sblock 128,64 --steps 256 --ring 64 > $OUT/sweep_q3d_pair_shapes.txt 2>&1
$S --envs quadrotor2d-v0,quadrotor2d-slungload-v0,quadrotor3d-slungload-v0 --vec 1 --n 131072,262144,524288,2097152,4194304 --fused 0,2 --steps 128 --ring 16 > $OUT/sweep_other_kinds_mid.txt 2>&1
unset RMV_FORCE_STREAM
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err
echo "bench rc=$?" >> $OUT/bench.err
ncu --set full --cache-control none --clock-control none --import-source on -k regex:k_step_stream_pair -s 2 -c 1 -f -o $OUT/pair_q3d_n20 \
    python tools/sweep.py --envs quadrotor3d-v0 --n 1048576 --vec 2 --layout aos --fused 0 --steps 16 --ring 16 > $OUT/ncu_a.log 2>&1
ncu -i $OUT/pair_q3d_n20.ncu-rep --page raw --csv > $OUT/pair_q3d_n20.raw.csv 2>/dev/null
ncu -i $OUT/pair_q3d_n20.ncu-rep --page details > $OUT/pair_q3d_n20.details.txt 2>/dev/null
ncu -i $OUT/pair_q3d_n20.ncu-rep --page source --csv --print-source sass > $OUT/pair_q3d_n20.sass.csv 2>/dev/null
rm -f $OUT/pair_q3d_n20.ncu-rep
