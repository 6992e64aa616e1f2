#!/bin/bash
# round-2 experiment Q: ptxas --register-usage-level 0 / 10 (the default is 5; 0-2 and 6-10 each give one other code shape)
cd "$(dirname "$0")/.."
run() { label=$1; shift; timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-trials 1e6 "$@" 2>&1 | tail -1 | python scripts/bench_value.py "$label"; }
cp startrackermpm_b200/libstmpm.so /tmp/libstmpm_keep.so
V=${VARIANTS:-base rul0 rul10}
for v in $V $V; do
  cp variants/libstmpm_$v.so startrackermpm_b200/libstmpm.so; touch startrackermpm_b200/libstmpm.so
  run "$v default"
done
for v in $V; do
  cp variants/libstmpm_$v.so startrackermpm_b200/libstmpm.so; touch startrackermpm_b200/libstmpm.so
  run "$v all_visible" --workload all_visible
done
cp /tmp/libstmpm_keep.so startrackermpm_b200/libstmpm.so; touch startrackermpm_b200/libstmpm.so
