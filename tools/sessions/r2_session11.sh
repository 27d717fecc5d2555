#!/bin/bash
OUT=gpurun_out/r2s11
mkdir -p $OUT
S="timeout 900 python tools/sweep.py --layout aos --vec 1"
$S --envs quadrotor3d-slungload-v0,quadrotor2d-slungload-v0,quadrotor3d-v0,quadrotor2d-v0 --n 65536,131072 --fused 0 --apre 0,1,-1 --steps 1024 --ring 64 > $OUT/sweep_apre_small.txt 2>&1
( time timeout 1800 python -m pytest tests -m gpu -x -q ) > $OUT/pytest.log 2>&1
echo "rc=$?" >> $OUT/pytest.log
timeout 300 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1
echo "smoke rc=$?" >> $OUT/smoke.log
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err
echo "bench rc=$?" >> $OUT/bench.err
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench_driver_flags.json 2> $OUT/bench_driver_flags.err
