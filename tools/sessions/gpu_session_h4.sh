#!/bin/bash
# four GPUs: sharded parity tests (incl. the one-process-per-GPU NCCL test) and the sharded bench with trace
OUT=gpurun_out/sh4
mkdir -p $OUT
( time timeout 300 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q ) > $OUT/pytest_sharded.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_sharded.log
CAPSB_TRACE=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus 4 --steps 2 --warmup 2 > $OUT/bench_4gpu.json 2> $OUT/bench_4gpu.err
tail -4 $OUT/pytest_sharded.log
tail -1 $OUT/bench_4gpu.json | cut -c1-1800
