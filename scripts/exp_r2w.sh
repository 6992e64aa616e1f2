#!/bin/bash
# round-2 experiment W: two stars per iteration chosen per instantiation (wFR: F = fp32 mode two-star, R = native RNG two-star)
cd "$(dirname "$0")/.."
cp startrackermpm_b200/libstmpm.so /tmp/libstmpm_keep.so
for v in ${VARIANTS:-w11 w00 w11 w00}; do
  cp variants/libstmpm_$v.so startrackermpm_b200/libstmpm.so; touch startrackermpm_b200/libstmpm.so
  echo "$v mode matrix:"; timeout 600 python scripts/mode_matrix.py 2>&1 | tail -1
done
cp /tmp/libstmpm_keep.so startrackermpm_b200/libstmpm.so; touch startrackermpm_b200/libstmpm.so
