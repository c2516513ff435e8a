#!/bin/bash
OUT=gpurun_out/sc
mkdir -p $OUT
for sl in 1 2 4; do echo "match slots $sl"; timeout 120 ./bin/radix_bench_slots$sl 4e8 10 0; done > $OUT/radix_bench.txt 2>&1
( time timeout 600 python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_gpu.log
CAPSB_TRACE=1 timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $OUT/trace.json 2> $OUT/trace.err
cat $OUT/radix_bench.txt; tail -3 $OUT/pytest_gpu.log; cat $OUT/trace.json | cut -c1-1500
