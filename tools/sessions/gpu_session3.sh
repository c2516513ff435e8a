#!/bin/bash
set -u
OUT=gpurun_out/s3
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -3 $OUT/pytest_gpu.log
for b in radix_bench radix_bench_SEQ_WRITE radix_bench_NO_BALLOT; do for mb in 2 3; do
  echo -n "$b mb=$mb: "; CAPSB_SCATTER_MIN_BLOCKS=$mb timeout 120 ./bin/$b 1e8 10 0
done; done 2>&1 | tee $OUT/radix_bench.txt
CAPSB_SCATTER_MIN_BLOCKS=2 timeout 300 ncu --set full --clock-control none --import-source on -k regex:radix_scatter -s 3 -c 2 -o $OUT/prof_scatter_v2_mb2 ./bin/radix_bench 1e8 2 0 > $OUT/ncu_mb2.log 2>&1
CAPSB_SCATTER_MIN_BLOCKS=3 timeout 300 ncu --set full --clock-control none --import-source on -k regex:radix_scatter -s 3 -c 2 -o $OUT/prof_scatter_v2_mb3 ./bin/radix_bench 1e8 2 0 > $OUT/ncu_mb3.log 2>&1
ls -la $OUT
