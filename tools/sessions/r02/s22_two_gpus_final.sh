#!/bin/bash
# round 2, session 22 (2 GPUs): the final tree through torchrun / NCCL, verified
OUT=gpurun_out/r02_s22
mkdir -p $OUT
( time timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline ) > $OUT/bench_genome3g_2gpu.json 2> $OUT/bench_genome3g_2gpu.err
echo "bench rc=$?" >> $OUT/bench_genome3g_2gpu.err
tail -3 $OUT/bench_genome3g_2gpu.err | cut -c1-300; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_s22/bench_genome3g_2gpu.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','stage_ms_rank0','verified','gpu_launches')}); print(d['e2e'])
PY
