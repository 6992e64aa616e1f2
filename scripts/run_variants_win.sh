#!/bin/bash
# scripts/run_variants_win.sh NAME...: each variants/libstmpm_NAME.so at every window in $WINS
cd "$(dirname "$0")/.."
cp startrackermpm_b200/libstmpm.so /tmp/libstmpm_keep.so
for v in "$@"; do
  cp variants/libstmpm_$v.so startrackermpm_b200/libstmpm.so
  for w in ${WINS:-128 512 2048}; do
    timeout 200 python bench.py --window $w --steps 3 --warmup 3 --no-cpu-baseline --e2e-trials 1e6 2>&1 | tail -1 | python scripts/bench_value.py "$v win=$w"
  done
done
cp /tmp/libstmpm_keep.so startrackermpm_b200/libstmpm.so
