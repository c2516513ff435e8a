#!/bin/bash
# round 2, session 13 (8 GPUs): NCCL parity at 8 ranks, bench at N=8 and N=4 (verified by rank 0)
OUT=gpurun_out/r02_s13
mkdir -p $OUT
nvidia-smi -L > $OUT/env.txt; nproc >> $OUT/env.txt; free -g >> $OUT/env.txt
( time timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29631 tools/sharded_check.py ) > $OUT/sharded_check_8.log 2>&1
echo "sharded_check rc=$?" >> $OUT/sharded_check_8.log; tail -14 $OUT/sharded_check_8.log | cut -c1-300
( time timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29633 bench.py --gpus 8 --steps 3 --warmup 3 ) > $OUT/bench_genome3g_8gpu.json 2> $OUT/bench_genome3g_8gpu.err
echo "bench8 rc=$?" >> $OUT/bench_genome3g_8gpu.err; tail -3 $OUT/bench_genome3g_8gpu.err | cut -c1-300
( time timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29635 bench.py --gpus 4 --steps 3 --warmup 3 ) > $OUT/bench_genome3g_4gpu.json 2> $OUT/bench_genome3g_4gpu.err
echo "bench4 rc=$?" >> $OUT/bench_genome3g_4gpu.err; tail -3 $OUT/bench_genome3g_4gpu.err | cut -c1-300
python - <<'PY'
import json
for g in (8,4):
    try:
        d=json.loads(open(f'gpurun_out/r02_s13/bench_genome3g_{g}gpu.json').read().strip().splitlines()[-1])
        print(g, {k:d[k] for k in ('value','ms_per_step','stage_ms_rank0','verified','gpu_launches')}); print(d['e2e']); print(d['config'])
    except Exception as e: print(g, 'no line', e)
PY
