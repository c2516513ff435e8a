#!/bin/bash
# round 2, session 25: L2 prefetch hint for the bucket a CTA of the local sort takes next
OUT=gpurun_out/r02_s25
mkdir -p $OUT
( time timeout 600 python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_gpu.log; tail -4 $OUT/pytest_gpu.log
( CAPSB_MSD_L2_PREFETCH=0 timeout 200 python bench.py --steps 3 --warmup 3 --no-e2e --no-verify --no-cpu-baseline ) > $OUT/bench_prefetch0.json 2> $OUT/bench_prefetch0.err
( timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline ) > $OUT/bench_verified.json 2> $OUT/bench_verified.err
echo "bench rc=$?" >> $OUT/bench_verified.err; tail -1 $OUT/bench_verified.err; python - <<'PY'
import json
for name in ('prefetch0', 'verified'):
    try:
        d=json.loads(open(f'gpurun_out/r02_s25/bench_{name}.json').read().strip().splitlines()[-1])
        print(name, round(d['ms_per_step'],2), d['stage_ms'], {k:round(v['ms_per_step'],2) for k,v in d['kernels'].items()}, d.get('verified',{}).get('code'), d.get('e2e',{}).get('ms_per_step'))
    except Exception as e:
        print(name, 'failed', e)
PY
