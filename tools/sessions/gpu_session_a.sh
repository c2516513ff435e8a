#!/bin/bash
# One GPU session: parity tests, bench lines, ncu launch list and a full capture of the dominant kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,driver_version --format=csv > gpurun_out/env.txt 2>&1
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
( time timeout 420 python bench.py --steps 3 --warmup 3 ) > gpurun_out/bench_genome3g.json 2> gpurun_out/bench_genome3g.err
( time timeout 120 python bench.py --workload random100m --steps 5 --warmup 3 ) > gpurun_out/bench_random100m.json 2> gpurun_out/bench_random100m.err
( time timeout 120 python bench.py --impl reference --steps 1 --warmup 0 ) > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_genome3g.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout 240 ncu --set full --clock-control none --import-source on -k regex:radix_scatter -s 4 -c 3 -o gpurun_out/scatter_full \
    python bench.py --workload random100m --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
