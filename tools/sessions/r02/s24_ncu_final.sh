#!/bin/bash
# round 2, session 24: ncu --set full of the final tree's local sort and of the first-pass / group-sort kernels of the refinement
OUT=gpurun_out/r02_s24
mkdir -p $OUT
( time timeout 600 ncu --set full --clock-control none --import-source on -k "regex:msd_local|key_lcp_count|tied_collect|group_sort" -c 6 -o $OUT/final_kernels_3g \
    python bench.py --steps 1 --warmup 0 --no-e2e --no-verify --no-cpu-baseline ) > $OUT/ncu_full.log 2>&1; tail -3 $OUT/ncu_full.log; ls -la $OUT
