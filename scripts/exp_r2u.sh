#!/bin/bash
# round-2 experiment U: two stars per star-loop iteration (STMPM_TWO_STARS) at 384 / 512 threads per CTA
cd "$(dirname "$0")/.."
run() { label=$1; shift; timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-trials 1e6 "$@" 2>&1 | tail -1 | python scripts/bench_value.py "$label"; }
cp startrackermpm_b200/libstmpm.so /tmp/libstmpm_keep.so
for v in ${VARIANTS:-base two384 two384c64 two512 tpb384 two384np base two384}; do
  cp variants/libstmpm_$v.so startrackermpm_b200/libstmpm.so; touch startrackermpm_b200/libstmpm.so
  run "$v default"
done
for v in ${VARIANTS_AV:-base two384 two384c64}; do
  cp variants/libstmpm_$v.so startrackermpm_b200/libstmpm.so; touch startrackermpm_b200/libstmpm.so
  run "$v all_visible" --workload all_visible
done
for v in ${VARIANTS_TEST-two384}; do
  cp variants/libstmpm_$v.so startrackermpm_b200/libstmpm.so; touch startrackermpm_b200/libstmpm.so
  echo "$v gpu tests:"; timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -12
done
cp /tmp/libstmpm_keep.so startrackermpm_b200/libstmpm.so; touch startrackermpm_b200/libstmpm.so
