#!/bin/bash
# round-2 experiment A: staging on/off, window sizes, dn clamp — device-resident bench value per variant
cd "$(dirname "$0")/.."
run() { label=$1; shift; timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-trials 1e6 "$@" 2>&1 | tail -1 | python scripts/bench_value.py "$label"; }
cp startrackermpm_b200/libstmpm.so /tmp/libstmpm_keep.so
cp variants/libstmpm_noclamp.so startrackermpm_b200/libstmpm.so; touch startrackermpm_b200/libstmpm.so
run "noclamp stage2048"
run "noclamp nostage256" --no-stage
run "noclamp stage1024" --window 1024
run "noclamp stage512" --window 512
run "noclamp stage256" --window 256
run "noclamp nostage2048" --no-stage --window 2048
cp variants/libstmpm_clamp.so startrackermpm_b200/libstmpm.so; touch startrackermpm_b200/libstmpm.so
run "clamp stage2048"
run "clamp nostage256" --no-stage
cp /tmp/libstmpm_keep.so startrackermpm_b200/libstmpm.so; touch startrackermpm_b200/libstmpm.so
run "main(footprint key) stage2048"
run "main(footprint key) nostage256" --no-stage
run "main disc-key stage2048" --key-radius-deg 9.99
python scripts/mode_matrix.py
