#!/bin/bash
# First session of the next round (two GPUs): validate the fused partition + exchange
# (CAPSB_SHARD_P2P=1) — thread transport on one device, then torchrun/NCCL with CUDA IPC —
# and time it against the default path.
#   gpurun --gpus 2 --timeout 600 -- 'bash tools/sessions/next_p2p_2gpu.sh'
OUT=gpurun_out/p2p
mkdir -p $OUT
( time CAPSB_TEST_P2P=1 timeout 300 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q -k "p2p" ) > $OUT/pytest_p2p.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_p2p.log
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 2 --warmup 2; }
run 29541 > $OUT/bench_default.json 2> $OUT/bench_default.err
CAPSB_SHARD_P2P=1 run 29542 > $OUT/bench_p2p.json 2> $OUT/bench_p2p.err
tail -4 $OUT/pytest_p2p.log
for f in default p2p; do tail -1 $OUT/bench_$f.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$f', round(d['ms_per_step'],2), d['stage_ms_rank0'], d['nvlink_bytes_per_step'])"; done
