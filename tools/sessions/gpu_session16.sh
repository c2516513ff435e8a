#!/bin/bash
set -u
OUT=gpurun_out/s16
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log
tail -12 $OUT/pytest.log
CAPSB_TRACE=1 timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e 2> $OUT/trace3g.err >/dev/null; grep -E "refine round|pair chains|stage" $OUT/trace3g.err | tail -34
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > $OUT/genome3g.json 2> $OUT/genome3g.err; python -c "
import json; d=json.loads(open('$OUT/genome3g.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['stage_ms'], d['roofline']['achieved'], d['roofline']['frac'])"
timeout 300 python bench.py --workload random100m --steps 5 --warmup 3 --no-cpu-baseline > $OUT/random100m.json 2>/dev/null; python -c "
import json; d=json.loads(open('$OUT/random100m.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['stage_ms'])"
