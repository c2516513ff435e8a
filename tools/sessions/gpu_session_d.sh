#!/bin/bash
OUT=gpurun_out/sd
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw,temperature.gpu --format=csv > $OUT/env.txt
for rep in 1 2; do for v in base base_carve merge merge_nocarve; do for k in 0 3; do echo -n "$v: "; timeout 60 ./bin/radix_bench_$v 4e8 10 0 $k | head -1; done; done; done > $OUT/radix_ab.txt 2>&1
( time timeout 600 python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_gpu.log
( time timeout 420 python bench.py --steps 3 --warmup 3 ) > $OUT/bench_genome3g.json 2> $OUT/bench_genome3g.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/launches_genome3g.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $OUT/ncu_launches.log 2>&1
cat $OUT/radix_ab.txt | cut -c1-230; tail -3 $OUT/pytest_gpu.log; cut -c1-900 $OUT/bench_genome3g.json
