#!/bin/bash
# round 2, session 10: sharded path with ranks found by the bucket owner; LARGE_IDX tool at small n; config 4 at 1 Gbp
OUT=gpurun_out/r02_s10
mkdir -p $OUT
( time CAPSB_TEST_P2P=1 timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q --durations=5 ) > $OUT/pytest_sharded.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_sharded.log
tail -8 $OUT/pytest_sharded.log
( time timeout 300 python tools/large_idx_check.py 3e7 2 ) > $OUT/large_idx_small.json 2> $OUT/large_idx_small.err; echo "rc=$?"; cut -c1-1500 $OUT/large_idx_small.json; tail -3 $OUT/large_idx_small.err
( time timeout 600 python tools/config4_check.py --text periodic ) > $OUT/config4_periodic_1g.json 2> $OUT/config4_periodic_1g.err; echo "rc=$?"; cat $OUT/config4_periodic_1g.json; tail -3 $OUT/config4_periodic_1g.err
( time timeout 900 python tools/config4_check.py --text fibonacci ) > $OUT/config4_fibonacci_1g.json 2> $OUT/config4_fibonacci_1g.err; echo "rc=$?"; cat $OUT/config4_fibonacci_1g.json; tail -3 $OUT/config4_fibonacci_1g.err
