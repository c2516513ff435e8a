#!/bin/bash
OUT=gpurun_out/sk
mkdir -p $OUT
( time timeout 600 python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_gpu.log
( time timeout 200 bash tools/test-correctness.sh 200000 2 ) > $OUT/test_correctness.log 2>&1
( time timeout 420 python bench.py --steps 3 --warmup 3 ) > $OUT/bench_genome3g.json 2> $OUT/bench_genome3g.err
tail -5 $OUT/pytest_gpu.log; cat $OUT/test_correctness.log; cut -c1-700 $OUT/bench_genome3g.json
