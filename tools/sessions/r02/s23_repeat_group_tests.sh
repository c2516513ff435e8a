#!/bin/bash
# round 2, session 23: the new parity cases (tied groups of 40 .. 5000 suffixes) on the final tree
OUT=gpurun_out/r02_s23
mkdir -p $OUT
( time timeout 600 python -m pytest tests -m gpu -k "repeat_groups" -q --durations=5 ) > $OUT/pytest_repeat_groups.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_repeat_groups.log; tail -12 $OUT/pytest_repeat_groups.log
CAPSB_TRACE=1 timeout 120 python - > $OUT/trace_repeat_groups.txt 2>&1 <<'PY'
import __graft_entry__ as g
pkg = g.load_package()
text = pkg.synth.repeat_groups(3_000_000)
sa = pkg.SuffixArray(text); sa.construct()
print({k: v for k, v in sa.stats().items() if k in ("refine_rounds", "tied_after_key_sort", "refine_counted", "refine_sorted", "msd_large_buckets", "msd_large_records", "pairs_chained")})
PY
grep -c "round:" $OUT/trace_repeat_groups.txt; tail -2 $OUT/trace_repeat_groups.txt
