#!/bin/bash
set -u
OUT=gpurun_out/s8
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log
tail -5 $OUT/pytest.log
timeout 900 python tools/quick_sharded.py 1e9 genome 1 2 > $OUT/quick_sharded_1g.txt 2>&1
grep -E "rank |single|ranks=" $OUT/quick_sharded_1g.txt | cut -c1-330
CAPSB_TRACE=1 timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > $OUT/genome3g.json 2> $OUT/genome3g.err
grep "refine round" $OUT/genome3g.err | tail -15
python - <<PY
import json
d=json.loads(open("$OUT/genome3g.json").read().strip().splitlines()[-1])
print("genome3g", "ms/step", round(d["ms_per_step"],3), "e2e ms", round(d["e2e"]["ms_per_step"],2), d["stage_ms"], d["roofline"]["achieved"])
PY
timeout 300 python bench.py --workload random100m --steps 5 --warmup 3 --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('random100m', d['ms_per_step'], d['e2e']['ms_per_step'], d['stage_ms'])"
