#!/bin/bash
# two GPUs: NCCL point-to-point channel count A/B, CPU affinity, end-to-end breakdown
OUT=gpurun_out/si2
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1; nproc >> $OUT/topo.txt; numactl -H >> $OUT/topo.txt 2>&1; df -h /dev/shm >> $OUT/topo.txt
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 2 --warmup 2; }
run 29521 > $OUT/bench_default.json 2> $OUT/bench_default.err
NCCL_MIN_P2P_NCHANNELS=16 run 29522 > $OUT/bench_p2p16.json 2> $OUT/bench_p2p16.err
NCCL_MIN_P2P_NCHANNELS=32 NCCL_MAX_P2P_NCHANNELS=32 run 29523 > $OUT/bench_p2p32.json 2> $OUT/bench_p2p32.err
for f in default p2p16 p2p32; do tail -1 $OUT/bench_$f.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$f', round(d['ms_per_step'],2), d['stage_ms_rank0'], d['e2e'])"; done
