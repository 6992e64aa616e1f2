#!/bin/bash
# round-2 experiment Z: run-time knobs on the committed v17 build: windows of 288 / 320 trials, footprint key table for pre-sampled launches
cd "$(dirname "$0")/.."
run() { label=$1; shift; timeout 100 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-trials 1e6 "$@" 2>&1 | tail -1 | python scripts/bench_value.py "$label"; }
run "v17 default"
run "v17 win=288" --window 288
run "v17 win=320" --window 320
run "v17 key-table=footprint" --key-table 1
run "v17 default"
