#!/bin/bash
# round 2, session 11 (2 GPUs): bench at N=2 with the owner-side rank search, LARGE_IDX at n = 2^32 + 1000, D2H probe
OUT=gpurun_out/r02_s11
mkdir -p $OUT
( time timeout 120 python tools/d2h_probe.py ) > $OUT/d2h_probe.json 2> $OUT/d2h_probe.err; cat $OUT/d2h_probe.json; tail -2 $OUT/d2h_probe.err
( time timeout 480 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 2 --steps 3 --warmup 2 ) > $OUT/bench_genome3g_2gpu.json 2> $OUT/bench_genome3g_2gpu.err
echo "bench rc=$?" >> $OUT/bench_genome3g_2gpu.err
tail -3 $OUT/bench_genome3g_2gpu.err | cut -c1-300; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_s11/bench_genome3g_2gpu.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','stage_ms_rank0','verified','gpu_launches')}); print(d['e2e'])
PY
( time timeout 900 python tools/large_idx_check.py ) > $OUT/large_idx_2p32.json 2> $OUT/large_idx_2p32.err; echo "large idx rc=$?"; cut -c1-2500 $OUT/large_idx_2p32.json; tail -5 $OUT/large_idx_2p32.err | cut -c1-400
