#!/bin/bash
# First GPU session of a round: environment facts, parity tests, bench lines, ncu evidence.
set -u
OUT=gpurun_out/s1
mkdir -p $OUT
{ nvidia-smi; nproc; free -g; lscpu | head -20; } > $OUT/env.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
timeout 600 python bench.py --workload random100m --steps 5 --warmup 3 > $OUT/bench_random100m.json 2> $OUT/bench_random100m.err
timeout 1200 python bench.py > $OUT/bench_genome3g.json 2> $OUT/bench_genome3g.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_random100m.csv \
    python bench.py --workload random100m --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:radix_scatter -s 10 -c 3 -o $OUT/prof_scatter \
    python bench.py --workload random100m --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
ls -la $OUT
tail -3 $OUT/pytest_gpu.log
cat $OUT/bench_random100m.json $OUT/bench_genome3g.json $OUT/bench_reference.json
tail -5 $OUT/bench_genome3g.err
