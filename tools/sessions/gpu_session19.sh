#!/bin/bash
# 2-GPU box: sharded path stage times (thread transport, 1 Gbp) and the torchrun bench at N=2 (3.1 Gbp)
set -u
OUT=gpurun_out/s19
mkdir -p $OUT
timeout 900 python tools/quick_sharded.py 1e9 genome 2 > $OUT/quick_sharded_1g.txt 2>&1
grep -E "rank |single|ranks=" $OUT/quick_sharded_1g.txt | cut -c1-400
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29751"
timeout 1200 $TR bench.py --gpus 2 --steps 2 --warmup 2 > $OUT/bench2_genome3g.json 2> $OUT/bench2_genome3g.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/s19/bench2_genome3g.json").read().strip().splitlines()[-1])
print("gpus", d["n_gpus"], "ms/step", round(d["ms_per_step"],2), "value", round(d["value"]/1e9,3), "G/s  e2e ms", round(d["e2e"]["ms_per_step"],2), d.get("stage_ms_rank0"), "nvlink", d.get("nvlink_bytes_per_step"), d["config"].get("shard_imbalance"))
PY
tail -3 $OUT/bench2_genome3g.err
