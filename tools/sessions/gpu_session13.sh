#!/bin/bash
# scatter layout variants (micro-benchmark) + parity tests + bench with the 512x8 layout
set -u
OUT=gpurun_out/s13
mkdir -p $OUT
bin/radix_bench 4e8 5 0 > $OUT/radix_bench.txt 2>&1
bin/radix_bench 4e8 5 32 >> $OUT/radix_bench.txt 2>&1
bin/radix_bench_SEQ_WRITE 4e8 5 0 1 >> $OUT/radix_bench.txt 2>&1
cat $OUT/radix_bench.txt
for c in 4 8 12; do CAPSB_SCATTER_CTAS_PER_SM=$c bin/radix_bench 4e8 5 0 1 | head -1; done | tee -a $OUT/radix_bench.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log
tail -4 $OUT/pytest.log
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > $OUT/genome3g.json 2> $OUT/genome3g.err; python -c "
import json; d=json.loads(open('$OUT/genome3g.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['stage_ms'], d['roofline']['achieved'], d['roofline']['frac'])"
