#!/bin/bash
set -u
OUT=gpurun_out/s6
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log
tail -5 $OUT/pytest.log
timeout 600 python tools/quick_sharded.py 1e8 genome 1 2 2>&1 | tee $OUT/quick_sharded_100m.txt
timeout 900 python tools/quick_sharded.py 1e9 genome 1 2>&1 | tee $OUT/quick_sharded_1g.txt
CAPSB_TRACE=1 timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/genome3g.json 2> $OUT/genome3g.err
grep capsb $OUT/genome3g.err | head -17
python - <<PY
import json
d=json.loads(open("$OUT/genome3g.json").read().strip().splitlines()[-1])
print("genome3g", "ms/step", round(d["ms_per_step"],3), "e2e ms", round(d["e2e"]["ms_per_step"],2), d["stage_ms"], d["config"].get("tied_after_key_sort"), d["config"].get("refine_rounds"))
PY
timeout 300 python bench.py --workload random100m --steps 5 --warmup 3 --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('random100m', d['ms_per_step'], d['stage_ms'])"
