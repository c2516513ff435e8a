#!/bin/bash
# round 2, session 5: ncu --set full of the MSD-sort kernels (v2) at 3.1 Gbp
OUT=gpurun_out/r02_s05
mkdir -p $OUT
( time timeout 1500 ncu --set full --clock-control none --import-source on -k "regex:msd_local|msd_scatter" -c 3 -o $OUT/msd_kernels_3g_v2 \
    python bench.py --steps 1 --warmup 0 --no-e2e --no-verify --no-cpu-baseline ) > $OUT/ncu_run.log 2>&1
echo "ncu rc=$?" >> $OUT/ncu_run.log
tail -3 $OUT/ncu_run.log
