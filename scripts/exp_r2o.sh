#!/bin/bash
# round-2 experiment O: hot instruction footprint (L1.5 instruction cache = 32 KB): single-copy sincos loop and friends
cd "$(dirname "$0")/.."
run() { label=$1; shift; timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-trials 1e6 "$@" 2>&1 | tail -1 | python scripts/bench_value.py "$label"; }
cp startrackermpm_b200/libstmpm.so /tmp/libstmpm_keep.so
V=${VARIANTS:-base scloop}
for v in $V $V; do
  cp variants/libstmpm_$v.so startrackermpm_b200/libstmpm.so; touch startrackermpm_b200/libstmpm.so
  run "$v default"
done
for v in $V; do
  cp variants/libstmpm_$v.so startrackermpm_b200/libstmpm.so; touch startrackermpm_b200/libstmpm.so
  run "$v all_visible" --workload all_visible
  echo "$v mode matrix:"; timeout 600 python scripts/mode_matrix.py 2>&1 | tail -1
done
cp /tmp/libstmpm_keep.so startrackermpm_b200/libstmpm.so; touch startrackermpm_b200/libstmpm.so
scripts/ncu_quick.sh $V
