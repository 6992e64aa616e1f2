#!/bin/bash
# round-2 experiment L: work-key tables stored level-major (a run touches a 41 KB / 330 KB slice instead of one sector per 128-byte row)
cd "$(dirname "$0")/.."
run() { label=$1; shift; timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-trials 1e6 "$@" 2>&1 | tail -1 | python scripts/bench_value.py "$label"; }
cp startrackermpm_b200/libstmpm.so /tmp/libstmpm_keep.so
for v in base keyT base keyT; do
  cp variants/libstmpm_$v.so startrackermpm_b200/libstmpm.so; touch startrackermpm_b200/libstmpm.so
  run "$v disc" --key-table 0
  run "$v footprint" --key-table 1
done
for v in base keyT; do
  cp variants/libstmpm_$v.so startrackermpm_b200/libstmpm.so; touch startrackermpm_b200/libstmpm.so
  for kt in 0 1; do echo "$v key_table=$kt mode matrix:"; MM_KEY_TABLE=$kt timeout 600 python scripts/mode_matrix.py 2>&1 | tail -1; done
done
cp variants/libstmpm_keyT.so startrackermpm_b200/libstmpm.so; touch startrackermpm_b200/libstmpm.so
run "keyT footprint all_visible" --workload all_visible --key-table 1
run "keyT disc all_visible" --workload all_visible --key-table 0
echo "keyT gpu tests:"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
cp /tmp/libstmpm_keep.so startrackermpm_b200/libstmpm.so; touch startrackermpm_b200/libstmpm.so
