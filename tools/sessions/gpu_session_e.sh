#!/bin/bash
OUT=gpurun_out/se
mkdir -p $OUT
( time timeout 600 python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_gpu.log
( time timeout 420 python bench.py --steps 3 --warmup 3 ) > $OUT/bench_genome3g.json 2> $OUT/bench_genome3g.err
( CAPSB_EARLY_SA=0 timeout 420 python bench.py --steps 3 --warmup 1 --no-cpu-baseline ) > $OUT/bench_genome3g_late_sa.json 2> $OUT/bench_genome3g_late_sa.err
CAPSB_TRACE=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $OUT/trace.json 2> $OUT/trace.err
tail -3 $OUT/pytest_gpu.log; cut -c1-900 $OUT/bench_genome3g.json
