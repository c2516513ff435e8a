#!/bin/bash
# round 2, session 8 (2 GPUs): sharded parity incl. NCCL transport and the fused partition pass, bench at N=2
OUT=gpurun_out/r02_s08
mkdir -p $OUT
nvidia-smi -L > $OUT/env.txt; nproc >> $OUT/env.txt; free -g >> $OUT/env.txt
( time CAPSB_TEST_P2P=1 timeout 1200 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q --durations=5 ) > $OUT/pytest_sharded.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_sharded.log
tail -15 $OUT/pytest_sharded.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 3 --warmup 2 ) > $OUT/bench_genome3g_2gpu.json 2> $OUT/bench_genome3g_2gpu.err
echo "bench rc=$?" >> $OUT/bench_genome3g_2gpu.err
tail -4 $OUT/bench_genome3g_2gpu.err | cut -c1-300; tail -1 $OUT/bench_genome3g_2gpu.json | cut -c1-3500
( time CAPSB_SHARD_P2P=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps 3 --warmup 2 ) > $OUT/bench_genome3g_2gpu_p2p.json 2> $OUT/bench_genome3g_2gpu_p2p.err
echo "bench rc=$?" >> $OUT/bench_genome3g_2gpu_p2p.err
tail -4 $OUT/bench_genome3g_2gpu_p2p.err | cut -c1-300; tail -1 $OUT/bench_genome3g_2gpu_p2p.json | cut -c1-2500
