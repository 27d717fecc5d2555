#!/bin/bash
OUT=gpurun_out/r2s9
mkdir -p $OUT
( time timeout 900 python -m pytest tests/test_gpu_vec_env.py tests/test_launcher_seam.py tests/test_vecnorm.py -m gpu -x -q ) > $OUT/pytest_vec.log 2>&1
echo "rc=$?" >> $OUT/pytest_vec.log
RMV_HOST_GEOMETRIC=0 timeout 300 python tools/e2e_multi_probe.py > $OUT/e2e_n1_uniform.jsonl 2> $OUT/e2e_a.err
RMV_HOST_GEOMETRIC=1 timeout 300 python tools/e2e_multi_probe.py > $OUT/e2e_n1_geometric.jsonl 2> $OUT/e2e_b.err
