#!/bin/bash
# round 2, session 17: group sort of the refinement in registers (bitonic, shuffles), and the size up
# to which the local sort orders a group by comparison instead of a second counting pass.
OUT=gpurun_out/r02_s17
mkdir -p $OUT
( time CAPSB_MSD_SMALL_GROUP=128 CAPSB_MSD_SMALL_GROUP_BIG=256 timeout 900 python -m pytest tests -m gpu -x -q --durations=5 ) > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_gpu.log; tail -10 $OUT/pytest_gpu.log
run() {  # name, env...
  name=$1; shift
  ( env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-verify --no-cpu-baseline ) > $OUT/bench_$name.json 2> $OUT/bench_$name.err
  python - $name <<'PY'
import json,sys
name=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/r02_s17/bench_{name}.json').read().strip().splitlines()[-1])
    print(name, round(d['ms_per_step'],2), d['stage_ms'], {k:round(v['ms_per_step'],2) for k,v in d['kernels'].items()})
except Exception as e:
    print(name, 'failed', e)
PY
}
run default CAPSB_X=0
run warp128 CAPSB_WARP_GROUP=128
run big256 CAPSB_MSD_SMALL_GROUP_BIG=256
run big256_small128 CAPSB_MSD_SMALL_GROUP_BIG=256 CAPSB_MSD_SMALL_GROUP=128
run big512_small256 CAPSB_MSD_SMALL_GROUP_BIG=512 CAPSB_MSD_SMALL_GROUP=256
run big128_small64 CAPSB_MSD_SMALL_GROUP_BIG=128 CAPSB_MSD_SMALL_GROUP=64
run big1024_small64 CAPSB_MSD_SMALL_GROUP_BIG=1024 CAPSB_MSD_SMALL_GROUP=64
# verified run (checker over all n entries) with the comparison limits raised
( time CAPSB_MSD_SMALL_GROUP=128 CAPSB_MSD_SMALL_GROUP_BIG=256 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline ) > $OUT/bench_verified.json 2> $OUT/bench_verified.err
echo "bench rc=$?" >> $OUT/bench_verified.err; tail -2 $OUT/bench_verified.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_s17/bench_verified.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','stage_ms','verified','gpu_launches')}); print(d['e2e'])
PY
