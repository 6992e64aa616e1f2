#!/bin/bash
# scripts/run_variants_modes.sh NAME...: scripts/mode_matrix.py (every mode x precision) for each variants/libstmpm_NAME.so
cd "$(dirname "$0")/.."
cp startrackermpm_b200/libstmpm.so /tmp/libstmpm_keep.so
for v in "$@"; do
  cp variants/libstmpm_$v.so startrackermpm_b200/libstmpm.so
  echo "$v $(timeout 200 python scripts/mode_matrix.py 2>/dev/null | tail -1)"
done
cp /tmp/libstmpm_keep.so startrackermpm_b200/libstmpm.so
