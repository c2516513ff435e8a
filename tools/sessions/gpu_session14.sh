#!/bin/bash
# pair-chain step: parity tests, trace, bench; smem primitive costs
set -u
OUT=gpurun_out/s14
mkdir -p $OUT
bin/smem_bench > $OUT/smem_bench.txt 2>&1; cat $OUT/smem_bench.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log
tail -12 $OUT/pytest.log
CAPSB_TRACE=1 timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e 2> $OUT/trace3g.err >/dev/null; grep -E "refine round|pairs finished" $OUT/trace3g.err | tail -24
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > $OUT/genome3g.json 2> $OUT/genome3g.err; python -c "
import json; d=json.loads(open('$OUT/genome3g.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['stage_ms'], d['roofline']['achieved'], d['roofline']['frac'])"
