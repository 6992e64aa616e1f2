#!/bin/bash
# round-2 experiment D: single persistent kernel (pipeline 0) vs two-kernel split pipeline (pipeline 1)
cd "$(dirname "$0")/.."
run() { label=$1; shift; timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-trials 1e6 "$@" 2>&1 | tail -1 | python scripts/bench_value.py "$label"; }
run "single kernel" --pipeline 0
run "split pipeline" --pipeline 1
run "single kernel" --pipeline 0
run "split pipeline" --pipeline 1
run "split pipeline all_visible (falls back)" --pipeline 1 --workload all_visible
run "auto all_visible" --workload all_visible
