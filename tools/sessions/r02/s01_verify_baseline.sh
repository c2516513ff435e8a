#!/bin/bash
# round 2, session 1: the round-1 kernels under the new tests and bench.py's post-run verification
OUT=gpurun_out/r02_s01
mkdir -p $OUT
nvidia-smi > $OUT/env.txt 2>&1; nproc >> $OUT/env.txt; free -g >> $OUT/env.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 ) > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_gpu.log
( time timeout 900 python bench.py --steps 3 --warmup 3 ) > $OUT/bench_genome3g.json 2> $OUT/bench_genome3g.err
echo "bench rc=$?" >> $OUT/bench_genome3g.err
tail -25 $OUT/pytest_gpu.log; tail -5 $OUT/bench_genome3g.err; cut -c1-3000 $OUT/bench_genome3g.json
