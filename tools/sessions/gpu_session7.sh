#!/bin/bash
set -u
OUT=gpurun_out/s7
mkdir -p $OUT
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/launches_genome3g.csv \
   python tools/quick_time.py 3.1e9 genome 1 > $OUT/ncu_quick.log 2>&1
tail -3 $OUT/ncu_quick.log
CAPSB_TRACE=1 timeout 900 python tools/quick_sharded.py 1e9 genome 1 > $OUT/quick_sharded_1g_trace.txt 2>&1
grep -E "capsb dev|rank 0|single" $OUT/quick_sharded_1g_trace.txt | grep -v "refine round" | head -80
