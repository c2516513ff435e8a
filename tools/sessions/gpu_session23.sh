#!/bin/bash
set -u
OUT=gpurun_out/s23
mkdir -p $OUT
bin/radix_bench 4e8 5 0 > $OUT/radix_bench.txt 2>&1; cat $OUT/radix_bench.txt
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log
tail -5 $OUT/pytest.log
CAPSB_TRACE=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e 2> $OUT/trace3g.err >/dev/null
awk '/refine: count/{c++} c==2' $OUT/trace3g.err | grep -vE "  round:" | cut -c1-110 | head -12
timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > $OUT/genome3g.json 2> $OUT/genome3g.err; python -c "
import json; d=json.loads(open('$OUT/genome3g.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['stage_ms'], d['roofline']['achieved'], d['roofline']['frac'])"
