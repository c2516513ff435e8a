#!/bin/bash
set -u
OUT=gpurun_out/s4
mkdir -p $OUT
for b in radix_bench radix_bench_SEQ_WRITE; do for mb in 2 3; do
  echo -n "$b mb=$mb: "; CAPSB_SCATTER_MIN_BLOCKS=$mb timeout 120 ./bin/$b 1e8 10 0
done; done 2>&1 | tee $OUT/radix_bench.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $OUT/pytest_parity.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_parity.log
tail -5 $OUT/pytest_parity.log
timeout 1500 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q > $OUT/pytest_sharded.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_sharded.log
tail -30 $OUT/pytest_sharded.log
CAPSB_TRACE=1 timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/genome3g_trace.json 2> $OUT/genome3g_trace.err
grep -m 40 capsb $OUT/genome3g_trace.err | head -40
python - <<PY
import json
d=json.loads(open("$OUT/genome3g_trace.json").read().strip().splitlines()[-1])
print("genome3g", "ms/step", round(d["ms_per_step"],3), "e2e ms", round(d["e2e"]["ms_per_step"],2), d["stage_ms"], d["config"].get("tied_after_key_sort"), d["config"].get("refine_rounds"))
PY
