#!/bin/bash
# 8-GPU box: the host-buffer ceiling at 2 / 4 / 8 ranks, then the contract bench at N = 8 (NCCL episode-stats path included)
OUT=gpurun_out/r2s5
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
lscpu | head -30 > $OUT/lscpu.txt
for n in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n tools/e2e_multi_probe.py > $OUT/e2e_probe_n$n.jsonl 2> $OUT/e2e_probe_n$n.err
done
timeout 600 python tools/e2e_multi_probe.py > $OUT/e2e_probe_n1.jsonl 2> $OUT/e2e_probe_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29600 bench.py --gpus 8 --steps 20 --warmup 5 > $OUT/bench_n8.json 2> $OUT/bench_n8.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/bench_n2.json 2> $OUT/bench_n2.err
