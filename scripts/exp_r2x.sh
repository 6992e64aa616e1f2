#!/bin/bash
# round-2 experiment X: fewer FP64-pipe instructions in the star loop (STMPM_TRIM64 = 1, 2) on the two-star build
cd "$(dirname "$0")/.."
run() { label=$1; shift; timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-trials 1e6 "$@" 2>&1 | tail -1 | python scripts/bench_value.py "$label"; }
cp startrackermpm_b200/libstmpm.so /tmp/libstmpm_keep.so
use() { cp variants/libstmpm_$1.so startrackermpm_b200/libstmpm.so; touch startrackermpm_b200/libstmpm.so; }
for v in ${VARIANTS:-x0 x1 x2 x0 x1 x2}; do use $v; run "$v default"; done
for v in ${VARIANTS_AV:-x0 x2}; do use $v; run "$v all_visible" --workload all_visible; done
for v in ${VARIANTS_TEST-x2}; do
  use $v; echo "$v mode matrix:"; timeout 600 python scripts/mode_matrix.py 2>&1 | tail -1
  echo "$v gpu tests:"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6
done
cp /tmp/libstmpm_keep.so startrackermpm_b200/libstmpm.so; touch startrackermpm_b200/libstmpm.so
