#!/bin/bash
# round 2, session 9: ranks of final suffixes found instead of published; 1024-thread local sort for big
# buckets; 6/12-slot local sort; rolling windows at level A.  Full GPU tests + bench + sanitizer on small cases.
OUT=gpurun_out/r02_s09
mkdir -p $OUT
( time timeout 900 python -m pytest tests -m gpu -x -q --durations=5 ) > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -12 $OUT/pytest_gpu.log
( time timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline ) > $OUT/bench_genome3g.json 2> $OUT/bench_genome3g.err
echo "bench rc=$?" >> $OUT/bench_genome3g.err
tail -3 $OUT/bench_genome3g.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_s09/bench_genome3g.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','stage_ms','verified','gpu_launches')}); print(d['e2e'])
for k,v in d['kernels'].items(): print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a not in('algorithmic_bytes','kernel')})
print(d['config']['key_sort'])
PY
( time timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden_fixture or tiny_inputs or (stage_key_sort and (acgt_70k or allA or acgt_6145 or polyA)) or (range_by_range and fibonacci and 7-True)" ) > $OUT/sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?" >> $OUT/sanitizer_memcheck.log; tail -6 $OUT/sanitizer_memcheck.log
( time timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "(stage_key_sort and (acgt_70k-False or allA_300k-False or acgt_6145-False)) or (golden_fixture and (fib or repeats))" ) > $OUT/sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?" >> $OUT/sanitizer_racecheck.log; tail -6 $OUT/sanitizer_racecheck.log
