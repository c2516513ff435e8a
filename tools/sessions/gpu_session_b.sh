#!/bin/bash
# Scatter-kernel variants (ballot ranking), parity, bench + trace.
OUT=gpurun_out/sb
mkdir -p $OUT
for n in 1e8 4e8; do
  timeout 120 ./bin/radix_bench $n 10 0
done > $OUT/radix_bench.txt 2>&1
timeout 60 ./bin/radix_bench_SEQ_WRITE 1e8 10 0 0 > $OUT/radix_bench_seq.txt 2>&1
for c in 4 8; do echo "ctas_per_sm=$c"; CAPSB_SCATTER_CTAS_PER_SM=$c timeout 60 ./bin/radix_bench 4e8 10 0 0; done > $OUT/radix_bench_grid.txt 2>&1
( time timeout 600 python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_gpu.log
( time timeout 420 python bench.py --steps 3 --warmup 3 ) > $OUT/bench_genome3g.json 2> $OUT/bench_genome3g.err
CAPSB_TRACE=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $OUT/trace.json 2> $OUT/trace.err
cat $OUT/radix_bench.txt $OUT/radix_bench_seq.txt $OUT/radix_bench_grid.txt; tail -3 $OUT/pytest_gpu.log
