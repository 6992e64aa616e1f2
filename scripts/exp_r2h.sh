#!/bin/bash
# round-2 experiment H: magnitude word kept alive by an immediate compare (no live register) vs the exact float test
cd "$(dirname "$0")/.."
run() { label=$1; shift; timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-trials 1e6 "$@" 2>&1 | tail -1 | python scripts/bench_value.py "$label"; }
cp startrackermpm_b200/libstmpm.so /tmp/libstmpm_keep.so
for v in magnan base k2arith magnan base k2arith; do
  cp variants/libstmpm_$v.so startrackermpm_b200/libstmpm.so; touch startrackermpm_b200/libstmpm.so
  run "$v default"
done
for v in magnan base k2arith; do
  cp variants/libstmpm_$v.so startrackermpm_b200/libstmpm.so; touch startrackermpm_b200/libstmpm.so
  echo "$v mode matrix:"; python scripts/mode_matrix.py
done
cp /tmp/libstmpm_keep.so startrackermpm_b200/libstmpm.so; touch startrackermpm_b200/libstmpm.so
