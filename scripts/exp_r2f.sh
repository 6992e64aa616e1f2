#!/bin/bash
# round-2 experiment F: effect of keeping the record's magnitude word alive (WAW fix) on every mode; split pipeline with dynamic groups
cd "$(dirname "$0")/.."
run() { label=$1; shift; timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-trials 1e6 "$@" 2>&1 | tail -1 | python scripts/bench_value.py "$label"; }
run "single kernel" --pipeline 0
run "split pipeline (dynamic groups)" --pipeline 1
python scripts/mode_matrix.py
cp startrackermpm_b200/libstmpm.so /tmp/libstmpm_keep.so
cp variants/libstmpm_park.so startrackermpm_b200/libstmpm.so; touch startrackermpm_b200/libstmpm.so
echo "previous build (before the magnitude-word fix):"; python scripts/mode_matrix.py
cp /tmp/libstmpm_keep.so startrackermpm_b200/libstmpm.so; touch startrackermpm_b200/libstmpm.so
