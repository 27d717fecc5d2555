#!/bin/bash
OUT=gpurun_out/r2s6
mkdir -p $OUT
( time timeout 1800 python -m pytest tests -m gpu -x -q ) > $OUT/pytest.log 2>&1
echo "rc=$?" >> $OUT/pytest.log
timeout 300 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1
echo "smoke rc=$?" >> $OUT/smoke.log
RMV_HOST_BATCH_SMALL=0 timeout 300 python tools/e2e_multi_probe.py > $OUT/e2e_n1_small_per_chunk.jsonl 2> $OUT/e2e_a.err
RMV_HOST_BATCH_SMALL=1 timeout 300 python tools/e2e_multi_probe.py > $OUT/e2e_n1_small_batched.jsonl 2> $OUT/e2e_b.err
bash tools/profile_round2.sh r2s6/prof > $OUT/profile.log 2>&1
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err
echo "bench rc=$?" >> $OUT/bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
