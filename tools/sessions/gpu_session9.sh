#!/bin/bash
# 2-GPU box: clean (no trace) bench at N=1 and N=2
set -u
OUT=gpurun_out/s9
mkdir -p $OUT
show() { python - "$1" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1].split('/')[-1], "gpus", d["n_gpus"], "ms/step", round(d["ms_per_step"],2), "value", round(d["value"]/1e9,3), "G/s  e2e ms", round(d["e2e"]["ms_per_step"],2), d.get("stage_ms") or d.get("stage_ms_rank0"), "scatter GB/s", round(d["roofline"]["achieved"] or 0,1), "nvlink", d.get("nvlink_bytes_per_step"))
PY
}
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench1_genome3g.json 2> $OUT/bench1_genome3g.err; show $OUT/bench1_genome3g.json
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29751"
timeout 1200 $TR bench.py --gpus 2 --steps 3 --warmup 3 > $OUT/bench2_genome3g.json 2> $OUT/bench2_genome3g.err; echo "rc=$?"; show $OUT/bench2_genome3g.json
tail -3 $OUT/bench2_genome3g.err
timeout 600 $TR bench.py --gpus 2 --workload genome100m --steps 3 --warmup 3 > $OUT/bench2_g100m.json 2> $OUT/bench2_g100m.err; show $OUT/bench2_g100m.json
timeout 600 python bench.py --workload genome100m --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench1_g100m.json 2>/dev/null; show $OUT/bench1_g100m.json
