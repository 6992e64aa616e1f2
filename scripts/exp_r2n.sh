#!/bin/bash
# round-2 experiment N: in-warp work donation at a warp-uniform point of the star loop (STMPM_REBALANCE)
cd "$(dirname "$0")/.."
run() { label=$1; shift; timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-trials 1e6 "$@" 2>&1 | tail -1 | python scripts/bench_value.py "$label"; }
cp startrackermpm_b200/libstmpm.so /tmp/libstmpm_keep.so
for v in ${VARIANTS:-base rebal base rebal}; do
  cp variants/libstmpm_$v.so startrackermpm_b200/libstmpm.so; touch startrackermpm_b200/libstmpm.so
  run "$v default"
done
for v in ${VARIANTS_AV:-base rebal}; do
  cp variants/libstmpm_$v.so startrackermpm_b200/libstmpm.so; touch startrackermpm_b200/libstmpm.so
  run "$v all_visible" --workload all_visible
done
for v in ${VARIANTS_MM:-rebal}; do
  cp variants/libstmpm_$v.so startrackermpm_b200/libstmpm.so; touch startrackermpm_b200/libstmpm.so
  echo "$v mode matrix:"; timeout 600 python scripts/mode_matrix.py 2>&1 | tail -1
  echo "$v gpu tests:"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -25
done
cp /tmp/libstmpm_keep.so startrackermpm_b200/libstmpm.so; touch startrackermpm_b200/libstmpm.so
