#!/bin/bash
# round 2, session 7: range-by-range construction with streamed results
OUT=gpurun_out/r02_s07
mkdir -p $OUT
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 ) > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -25 $OUT/pytest_gpu.log
( time timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline ) > $OUT/bench_genome3g.json 2> $OUT/bench_genome3g.err
echo "bench rc=$?" >> $OUT/bench_genome3g.err
tail -3 $OUT/bench_genome3g.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_s07/bench_genome3g.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','stage_ms','verified','gpu_launches')}); print(d['e2e'])
for k,v in d['kernels'].items(): print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a not in('algorithmic_bytes','kernel')})
PY
( CAPSB_TRACE=1 timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-verify ) > $OUT/trace_bench.json 2> $OUT/trace_genome3g.txt
grep -c . $OUT/trace_genome3g.txt; tail -150 $OUT/trace_genome3g.txt | cut -c1-160
