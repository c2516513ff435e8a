#!/bin/bash
# ncu --set full of the scatter kernel (micro-benchmark, 400 M pairs) + launch list of one genome3g construction
set -u
OUT=gpurun_out/s12
mkdir -p $OUT
bin/radix_bench 4e8 5 0 > $OUT/radix_bench.txt 2>&1; cat $OUT/radix_bench.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:radix_scatter -s 3 -c 2 -o $OUT/scatter_v3 bin/radix_bench 4e8 2 0 > $OUT/ncu_scatter.log 2>&1; tail -3 $OUT/ncu_scatter.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_genome3g.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > $OUT/bench_under_ncu.log 2>&1; tail -2 $OUT/bench_under_ncu.log
