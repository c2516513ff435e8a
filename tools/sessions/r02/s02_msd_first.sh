#!/bin/bash
# round 2, session 2: first run of the packed-record MSD key sort
OUT=gpurun_out/r02_s02
mkdir -p $OUT
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stage_key_sort" ) > $OUT/pytest_keysort.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_keysort.log
tail -30 $OUT/pytest_keysort.log
( timeout 300 python tools/key_sort_time.py 1e8 3 ) > $OUT/keysort_100m_msd.txt 2>&1; cat $OUT/keysort_100m_msd.txt
( timeout 300 python tools/key_sort_time.py 1e8 2 lsd ) > $OUT/keysort_100m_lsd.txt 2>&1; cat $OUT/keysort_100m_lsd.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 ) > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -30 $OUT/pytest_gpu.log
( time timeout 900 python bench.py --steps 3 --warmup 3 ) > $OUT/bench_genome3g.json 2> $OUT/bench_genome3g.err
echo "bench rc=$?" >> $OUT/bench_genome3g.err
tail -5 $OUT/bench_genome3g.err; cut -c1-2500 $OUT/bench_genome3g.json
