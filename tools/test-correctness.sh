#!/bin/bash
# Random-seed check of the GPU CLI against the unmodified reference CLI (what the reference's
# utils/test-correctness.sh does against binaries that are not in its repository):
#   tools/test-correctness.sh [N] [rounds]
# generates a random ACGT string of length N with a fresh seed, runs bin/caps_sa and
# oracle/_ref/caps_sa_ref on it and compares the dump files byte for byte.
set -u
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
N=${1:-1000}
ROUNDS=${2:-1}
GPU="$ROOT/bin/caps_sa"
REF="$ROOT/oracle/_ref/caps_sa_ref"
[ -x "$GPU" ] || { echo "missing $GPU (python __graft_entry__.py)"; exit 2; }
[ -x "$REF" ] || { echo "missing $REF (make -C oracle ref)"; exit 2; }
[ "$N" -ge 64 ] || { echo "N must be at least 64 (the reference divides by zero below 16)"; exit 2; }
P=$(( N / 64 < 2 ? 2 : (N / 64 > 256 ? 256 : N / 64) ))
status=0
for r in $(seq 1 "$ROUNDS"); do
  seed=$RANDOM
  tmp=$(mktemp -d)
  python3 - "$seed" "$N" > "$tmp/in.txt" <<'PY'
import random, sys
random.seed(int(sys.argv[1]))
print("".join(random.choice("ACGT") for _ in range(int(sys.argv[2]))))
PY
  t0=$(date +%s.%N); "$REF" "$tmp/in.txt" "$tmp/ref.bin" "$P" 2> "$tmp/ref.err"; t1=$(date +%s.%N)
  "$GPU" "$tmp/in.txt" "$tmp/gpu.bin" "$P" 2> "$tmp/gpu.err"; t2=$(date +%s.%N)
  if cmp -s "$tmp/ref.bin" "$tmp/gpu.bin"; then result="\033[32mcorrect\033[0m"; else result="\033[31mincorrect\033[0m"; status=1; fi
  echo "Random seed: $seed"
  echo "True program runtime: $(awk "BEGIN{printf \"%.3f\", $t1 - $t0}") seconds"
  echo "Test program runtime: $(awk "BEGIN{printf \"%.3f\", $t2 - $t1}") seconds"
  echo -e "Output correctness: $result"
  rm -rf "$tmp"
done
exit $status
