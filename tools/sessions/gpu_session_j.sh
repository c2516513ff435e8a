#!/bin/bash
# Final single-GPU session of the round: parity tests, bench lines, launch list, full ncu captures,
# config 4 at full size.
OUT=gpurun_out/sj
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,driver_version --format=csv > $OUT/env.txt 2>&1
( time timeout 600 python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_gpu.log
( time timeout 420 python bench.py --steps 3 --warmup 3 ) > $OUT/bench_genome3g.json 2> $OUT/bench_genome3g.err
( time timeout 120 python bench.py --workload random100m --steps 5 --warmup 3 ) > $OUT/bench_random100m.json 2> $OUT/bench_random100m.err
( time timeout 120 python bench.py --impl reference --steps 1 --warmup 0 ) > $OUT/bench_reference.json 2> $OUT/bench_reference.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/launches_genome3g.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $OUT/ncu_launches.log 2>&1
timeout 240 ncu --set full --clock-control none --import-source on -k regex:"radix_scatter|key_lcp_count|radix_hist" -c 12 -o $OUT/kernels_full \
    python bench.py --workload random100m --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > $OUT/ncu_full.log 2>&1
( time timeout 400 python tools/config4_check.py --n 1e9 --text periodic ) > $OUT/config4_periodic.json 2> $OUT/config4_periodic.err
tail -3 $OUT/pytest_gpu.log; cut -c1-600 $OUT/bench_genome3g.json; cat $OUT/config4_periodic.json; tail -5 $OUT/config4_periodic.err
