#!/bin/bash
# round-2 experiment B: key tables per mode, staging with streaming hints, window sizes around 256
cd "$(dirname "$0")/.."
run() { label=$1; shift; timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-trials 1e6 "$@" 2>&1 | tail -1 | python scripts/bench_value.py "$label"; }
cp startrackermpm_b200/libstmpm.so /tmp/libstmpm_keep.so
cp variants/libstmpm_main.so startrackermpm_b200/libstmpm.so; touch startrackermpm_b200/libstmpm.so
run "main default (disc, 256, gather)"
run "main footprint-key 256" --key-table 1
run "main win320" --window 320
run "main win384" --window 384
run "main stage2048 footprint" --stage --key-table 1
run "main stage512 disc" --stage --window 512
cp variants/libstmpm_stagecs.so startrackermpm_b200/libstmpm.so; touch startrackermpm_b200/libstmpm.so
run "stagecs stage2048 footprint" --stage --key-table 1
run "stagecs stage2048 disc" --stage
run "stagecs stage1024 footprint" --stage --key-table 1 --window 1024
cp /tmp/libstmpm_keep.so startrackermpm_b200/libstmpm.so; touch startrackermpm_b200/libstmpm.so
python scripts/mode_matrix.py
