#!/bin/bash
# round 2, session 3: ncu --set full of the five MSD-sort kernels at 3.1 Gbp (one construction)
OUT=gpurun_out/r02_s03b
mkdir -p $OUT
( time timeout 1500 ncu --set full --clock-control none --import-source on -k "regex:msd_local|msd_scatter|msd_hist" -s 2 -c 3 -o $OUT/msd_kernels_3g_b \
    python bench.py --steps 1 --warmup 0 --no-e2e --no-verify --no-cpu-baseline ) > $OUT/ncu_run.log 2>&1
echo "ncu rc=$?" >> $OUT/ncu_run.log
tail -5 $OUT/ncu_run.log
ls -la $OUT
