#!/bin/bash
# round 2, session 12: local-sort A/B (counters per record), CLI wall times (configs 1, 2), config 2 bench line
OUT=gpurun_out/r02_s12
mkdir -p $OUT
( timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stage_key_sort or cli or fused" ) > $OUT/pytest_quick.log 2>&1; tail -3 $OUT/pytest_quick.log
for extra in 1 0; do
  ( CAPSB_MSD_LOCAL_EXTRA=$extra timeout 300 python bench.py --steps 3 --warmup 1 --no-e2e --no-verify --no-cpu-baseline ) > $OUT/bench_local_extra$extra.json 2> $OUT/bench_local_extra$extra.err
  python - <<PY
import json
d=json.loads(open('$OUT/bench_local_extra$extra.json').read().strip().splitlines()[-1])
print('extra=$extra', d['ms_per_step'], d['stage_ms'], {k:round(v['ms_per_step'],2) for k,v in d['kernels'].items()})
PY
done
( timeout 300 python tools/cli_time.py config1 ) > $OUT/cli_config1.json 2> $OUT/cli_config1.err; cut -c1-1800 $OUT/cli_config1.json
( timeout 600 python tools/cli_time.py config2 ) > $OUT/cli_config2.json 2> $OUT/cli_config2.err; cut -c1-1800 $OUT/cli_config2.json
( timeout 600 python tools/cli_time.py config2 256 ) > $OUT/cli_config2_p256.json 2> $OUT/cli_config2_p256.err; cut -c1-1800 $OUT/cli_config2_p256.json
( timeout 300 python bench.py --workload random100m --steps 5 --warmup 3 ) > $OUT/bench_random100m.json 2> $OUT/bench_random100m.err; cut -c1-1500 $OUT/bench_random100m.json
