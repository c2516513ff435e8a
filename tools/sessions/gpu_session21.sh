#!/bin/bash
# 2-GPU box: torchrun bench at N=2, small workload first (fail fast), then 3.1 Gbp
set -u
OUT=gpurun_out/s21
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29751"
show() { python - "$1" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("gpus", d["n_gpus"], "ms/step", round(d["ms_per_step"],2), "value", round(d["value"]/1e9,3), "G/s  e2e", d["e2e"], d.get("stage_ms_rank0"), "nvlink", d.get("nvlink_bytes_per_step"), d["config"].get("shard_imbalance"))
PY
}
timeout 150 $TR bench.py --gpus 2 --workload genome100m --steps 3 --warmup 3 > $OUT/bench2_g100m.json 2> $OUT/bench2_g100m.err; rc=$?; echo "small rc=$rc"
if [ $rc -ne 0 ]; then grep -E "Error|error|rank" $OUT/bench2_g100m.err | head -20; exit 1; fi
show $OUT/bench2_g100m.json
timeout 400 $TR bench.py --gpus 2 --steps 3 --warmup 3 > $OUT/bench2_genome3g.json 2> $OUT/bench2_genome3g.err; echo "rc=$?"
show $OUT/bench2_genome3g.json
grep -E "Error|error" $OUT/bench2_genome3g.err | head -5
