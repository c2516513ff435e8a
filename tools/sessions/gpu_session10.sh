#!/bin/bash
set -u
OUT=gpurun_out/s10
mkdir -p $OUT
show() { python - "$1" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1].split('/')[-1], "gpus", d["n_gpus"], "ms/step", round(d["ms_per_step"],2), "value", round(d["value"]/1e9,3), "G/s  e2e ms", round(d["e2e"]["ms_per_step"],2), d.get("stage_ms") or d.get("stage_ms_rank0"), "scatter GB/s", round(d["roofline"]["achieved"] or 0,1))
PY
}
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log
tail -15 $OUT/pytest.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/genome3g.json 2> $OUT/genome3g.err; show $OUT/genome3g.json
CAPSB_TRACE=1 timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline 2>&1 >/dev/null | grep "refine round" | head -16
timeout 300 python bench.py --workload random100m --steps 5 --warmup 3 --no-cpu-baseline > $OUT/random100m.json 2>/dev/null; show $OUT/random100m.json
timeout 900 python tools/quick_sharded.py 1e9 genome 1 2 > $OUT/quick_sharded_1g.txt 2>&1
grep -E "rank |single|ranks=" $OUT/quick_sharded_1g.txt | cut -c1-330
