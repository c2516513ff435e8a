#!/bin/bash
# 2-GPU box: parity (both shard modes, NCCL check), then the torchrun bench at N=2 (3.1 Gbp) and N=1
set -u
OUT=gpurun_out/s20
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log
tail -12 $OUT/pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29751"
timeout 1200 $TR bench.py --gpus 2 --steps 2 --warmup 2 > $OUT/bench2_genome3g.json 2> $OUT/bench2_genome3g.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/s20/bench2_genome3g.json").read().strip().splitlines()[-1])
print("gpus", d["n_gpus"], "ms/step", round(d["ms_per_step"],2), "value", round(d["value"]/1e9,3), "G/s  e2e ms", round(d["e2e"]["ms_per_step"],2), d["e2e"], d.get("stage_ms_rank0"), "nvlink", d.get("nvlink_bytes_per_step"), d["config"].get("shard_imbalance"))
PY
tail -5 $OUT/bench2_genome3g.err
