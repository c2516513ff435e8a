#!/bin/bash
# validates HEAD after the container re-creation: GPU tests, both bench workloads, a trace of the refinement rounds
set -u
OUT=gpurun_out/s11
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $OUT/env.txt; nproc >> $OUT/env.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log
tail -8 $OUT/pytest.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/genome3g.json 2> $OUT/genome3g.err; cat $OUT/genome3g.json
CAPSB_TRACE=1 timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline 2> $OUT/trace3g.err >/dev/null; grep -E "refine round|stage|ms" $OUT/trace3g.err | head -60
timeout 300 python bench.py --workload random100m --steps 5 --warmup 3 --no-cpu-baseline > $OUT/random100m.json 2>/dev/null; cat $OUT/random100m.json
