#!/bin/bash
OUT=gpurun_out/sl
mkdir -p $OUT
( time timeout 300 python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_gpu.log
( time timeout 200 python bench.py --steps 3 --warmup 2 ) > $OUT/bench_genome3g.json 2> $OUT/bench_genome3g.err
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/launches_genome3g.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $OUT/ncu_launches.log 2>&1
tail -5 $OUT/pytest_gpu.log; cut -c1-900 $OUT/bench_genome3g.json
