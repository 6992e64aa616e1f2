#!/bin/bash
# round-2 experiment K: 8-byte shared-memory catalog entries (STMPM_CAT8: 73 KB instead of 146 KB -> 164 KB carve-out, 92 KB L1)
cd "$(dirname "$0")/.."
run() { label=$1; shift; timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-trials 1e6 "$@" 2>&1 | tail -1 | python scripts/bench_value.py "$label"; }
cp startrackermpm_b200/libstmpm.so /tmp/libstmpm_keep.so
for v in base cat8 cap32 base cat8; do
  cp variants/libstmpm_$v.so startrackermpm_b200/libstmpm.so; touch startrackermpm_b200/libstmpm.so
  run "$v default"
done
for v in base cat8; do
  cp variants/libstmpm_$v.so startrackermpm_b200/libstmpm.so; touch startrackermpm_b200/libstmpm.so
  run "$v all_visible" --workload all_visible
done
cp variants/libstmpm_cat8.so startrackermpm_b200/libstmpm.so; touch startrackermpm_b200/libstmpm.so
echo "cat8 mode matrix:"; timeout 600 python scripts/mode_matrix.py 2>&1 | tail -1
echo "cat8 gpu tests:"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
cp /tmp/libstmpm_keep.so startrackermpm_b200/libstmpm.so; touch startrackermpm_b200/libstmpm.so
