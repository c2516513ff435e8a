#!/bin/bash
# profiles of the current state: ncu --set full of the product scatter kernel, launch list of one 3.1 Gbp construction
set -u
OUT=gpurun_out/s25
mkdir -p $OUT
timeout 200 ncu --set full --clock-control none --import-source on -k regex:radix_scatter -s 3 -c 2 -o $OUT/scatter_v4 bin/radix_bench 4e8 2 0 0 > $OUT/ncu_scatter.log 2>&1; tail -2 $OUT/ncu_scatter.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_genome3g.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > $OUT/bench_under_ncu.log 2>&1; tail -1 $OUT/bench_under_ncu.log | cut -c1-200
