#!/bin/bash
# Scatter-kernel variants + adaptive key width: parity, then timing.
set -u
OUT=gpurun_out/s2
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -3 $OUT/pytest_gpu.log
for mb in 2 3 4; do for per in 6 12; do
  CAPSB_SCATTER_MIN_BLOCKS=$mb CAPSB_SCATTER_CTAS_PER_SM=$per timeout 300 python bench.py --workload random100m --steps 5 --warmup 3 --no-cpu-baseline > $OUT/r100m_mb${mb}_per${per}.json 2>&1
  python - <<PY
import json
d=json.loads(open("$OUT/r100m_mb${mb}_per${per}.json").read().strip().splitlines()[-1])
print("mb=$mb per=$per", "ms/step", round(d["ms_per_step"],3), "scatter avg ms", round(d["roofline"]["avg_launch_ms"],4), "GB/s", round(d["roofline"]["achieved"],1), d["stage_ms"])
PY
done; done
CAPSB_KEY_BITS=64 timeout 300 python bench.py --workload random100m --steps 5 --warmup 3 --no-cpu-baseline > $OUT/r100m_key64.json 2>&1
CAPSB_KEY_BITS=32 timeout 300 python bench.py --workload random100m --steps 5 --warmup 3 --no-cpu-baseline > $OUT/r100m_key32.json 2>&1
timeout 300 python bench.py --workload genome100m --steps 5 --warmup 3 --no-cpu-baseline > $OUT/g100m.json 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/genome3g.json 2>&1
for f in r100m_key64 r100m_key32 g100m genome3g; do python - <<PY
import json
d=json.loads(open("$OUT/$f.json").read().strip().splitlines()[-1])
print("$f", "ms/step", round(d["ms_per_step"],3), "e2e ms", round(d["e2e"]["ms_per_step"],2), "scatter GB/s", round(d["roofline"]["achieved"],1), d["stage_ms"], d["config"].get("tied_after_key_sort"), d["config"].get("refine_rounds"))
PY
done
