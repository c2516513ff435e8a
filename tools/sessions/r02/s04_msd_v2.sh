#!/bin/bash
# round 2, session 4: MSD sort v2 (sector-aligned flushes with carry, blocked text loads, split histogram,
# padded counters in the local sort)
OUT=gpurun_out/r02_s04
mkdir -p $OUT
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stage_key_sort" ) > $OUT/pytest_keysort.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_keysort.log
tail -5 $OUT/pytest_keysort.log
( timeout 300 python tools/key_sort_time.py 1e8 3 ) > $OUT/keysort_100m_msd.txt 2>&1; cat $OUT/keysort_100m_msd.txt
( time timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline ) > $OUT/bench_genome3g.json 2> $OUT/bench_genome3g.err
echo "bench rc=$?" >> $OUT/bench_genome3g.err
tail -3 $OUT/bench_genome3g.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_s04/bench_genome3g.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','stage_ms','verified')}); print(d['e2e'])
for k,v in d['kernels'].items(): print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a not in('algorithmic_bytes','kernel')})
print(d['config']['key_sort'])
PY
( time timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "regex:msd_local|msd_scatter|msd_hist" -c 5 --csv --log-file $OUT/msd_dram_3g.csv \
    python bench.py --steps 1 --warmup 0 --no-e2e --no-verify --no-cpu-baseline ) > $OUT/ncu_run.log 2>&1
cat $OUT/msd_dram_3g.csv | cut -c1-400 | tail -20
