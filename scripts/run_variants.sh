#!/bin/bash
# scripts/run_variants.sh NAME...: on the GPU box, times bench.py's device-resident value for each variants/libstmpm_NAME.so
cd "$(dirname "$0")/.."
cp startrackermpm_b200/libstmpm.so /tmp/libstmpm_keep.so
for v in "$@"; do
  cp variants/libstmpm_$v.so startrackermpm_b200/libstmpm.so
  timeout 200 python bench.py --steps ${STEPS:-5} --warmup 3 --no-cpu-baseline --e2e-trials 1e6 2>&1 | tail -1 | python scripts/bench_value.py "$v"
done
cp /tmp/libstmpm_keep.so startrackermpm_b200/libstmpm.so
