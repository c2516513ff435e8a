#!/bin/bash
# round 2, session 14: pair-chain order A/B, full GPU tests, verified bench, ncu launch list and full captures of the final kernels
OUT=gpurun_out/r02_s14
mkdir -p $OUT
for pf in 1 0; do
  ( CAPSB_PAIRS_FIRST=$pf timeout 300 python bench.py --steps 3 --warmup 1 --no-e2e --no-verify --no-cpu-baseline ) > $OUT/bench_pairs_first$pf.json 2> $OUT/bench_pairs_first$pf.err
  python - <<PY
import json
d=json.loads(open('$OUT/bench_pairs_first$pf.json').read().strip().splitlines()[-1])
print('pairs_first=$pf', d['ms_per_step'], d['stage_ms'], d['gpu_launches'])
PY
done
( time timeout 900 python -m pytest tests -m gpu -x -q --durations=5 ) > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_gpu.log; tail -10 $OUT/pytest_gpu.log
( time timeout 600 python bench.py --steps 3 --warmup 3 ) > $OUT/bench_genome3g.json 2> $OUT/bench_genome3g.err
echo "bench rc=$?" >> $OUT/bench_genome3g.err; tail -2 $OUT/bench_genome3g.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_s14/bench_genome3g.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','stage_ms','verified','gpu_launches')}); print(d['e2e']); print(d['roofline']); print(d['cpu_baseline'])
PY
( time timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $OUT/launches_genome3g.csv \
    python bench.py --steps 1 --warmup 0 --no-e2e --no-verify --no-cpu-baseline ) > $OUT/ncu_launches.log 2>&1; wc -l $OUT/launches_genome3g.csv
( time timeout 900 ncu --set full --clock-control none --import-source on -k "regex:msd_local|msd_scatter" -c 4 -o $OUT/msd_kernels_3g_final \
    python bench.py --steps 1 --warmup 0 --no-e2e --no-verify --no-cpu-baseline ) > $OUT/ncu_full.log 2>&1; tail -2 $OUT/ncu_full.log
