#!/bin/bash
# round 2, session 21: level A shifts a thread's windows out of the first one (40-bit keys of DNA)
OUT=gpurun_out/r02_s21
mkdir -p $OUT
( time timeout 900 python -m pytest tests -m gpu -x -q --durations=5 ) > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_gpu.log; tail -6 $OUT/pytest_gpu.log
run() {  # name, env...
  name=$1; shift
  ( env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-verify --no-cpu-baseline ) > $OUT/bench_$name.json 2> $OUT/bench_$name.err
  python - $name <<'PY'
import json,sys
name=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/r02_s21/bench_{name}.json').read().strip().splitlines()[-1])
    print(name, round(d['ms_per_step'],2), d['stage_ms'], {k:round(v['ms_per_step'],2) for k,v in d['kernels'].items()})
except Exception as e:
    print(name, 'failed', e)
PY
}
run short0 CAPSB_MSD_SHORT=0
run short1 CAPSB_MSD_SHORT=1
( time timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline ) > $OUT/bench_verified.json 2> $OUT/bench_verified.err
echo "bench rc=$?" >> $OUT/bench_verified.err; tail -2 $OUT/bench_verified.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_s21/bench_verified.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','stage_ms','verified','gpu_launches')}); print(d['e2e'])
print({k:round(v['ms_per_step'],2) for k,v in d['kernels'].items()})
PY
