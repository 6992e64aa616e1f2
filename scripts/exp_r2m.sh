#!/bin/bash
# round-2 experiment M: 384 threads per CTA (12 warps, 168 registers, no spills) with the register-hungry variants on top
cd "$(dirname "$0")/.."
run() { label=$1; shift; timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-trials 1e6 "$@" 2>&1 | tail -1 | python scripts/bench_value.py "$label"; }
cp startrackermpm_b200/libstmpm.so /tmp/libstmpm_keep.so
for v in base tpb384 tpb384_nopark tpb384_idx tpb384_pp tpb384_cap64 tpb256 base tpb384; do
  cp variants/libstmpm_$v.so startrackermpm_b200/libstmpm.so; touch startrackermpm_b200/libstmpm.so
  run "$v default"
done
for v in base tpb384 tpb384_cap64 tpb256; do
  cp variants/libstmpm_$v.so startrackermpm_b200/libstmpm.so; touch startrackermpm_b200/libstmpm.so
  run "$v all_visible" --workload all_visible
done
cp /tmp/libstmpm_keep.so startrackermpm_b200/libstmpm.so; touch startrackermpm_b200/libstmpm.so
