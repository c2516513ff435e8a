#!/bin/bash
# round 2, session 16: mid-size groups of the local sort by warps; final tree — full GPU tests, verified bench, smoke
OUT=gpurun_out/r02_s16
mkdir -p $OUT
( time timeout 900 python -m pytest tests -m gpu -x -q --durations=5 ) > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_gpu.log; tail -10 $OUT/pytest_gpu.log
( time timeout 120 python -c "import __graft_entry__ as g; g.smoke()" ) > $OUT/smoke.log 2>&1; tail -3 $OUT/smoke.log
( time timeout 600 python bench.py --steps 5 --warmup 3 ) > $OUT/bench_genome3g.json 2> $OUT/bench_genome3g.err
echo "bench rc=$?" >> $OUT/bench_genome3g.err; tail -2 $OUT/bench_genome3g.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_s16/bench_genome3g.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','stage_ms','verified','gpu_launches')}); print(d['e2e']); print(d['roofline'])
for k,v in d['kernels'].items(): print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a not in('algorithmic_bytes','kernel')})
PY
( time timeout 300 python bench.py --impl reference --steps 2 --warmup 1 ) > $OUT/bench_reference.json 2> $OUT/bench_reference.err; cut -c1-900 $OUT/bench_reference.json
