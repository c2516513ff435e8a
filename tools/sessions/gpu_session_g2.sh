#!/bin/bash
# two GPUs: sharded bench with the per-stage trace of both ranks
OUT=gpurun_out/sg2
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
CAPSB_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 2 --warmup 2 > $OUT/bench_2gpu.json 2> $OUT/bench_2gpu.err
tail -1 $OUT/bench_2gpu.json | cut -c1-1800
grep -c capsb $OUT/bench_2gpu.err
