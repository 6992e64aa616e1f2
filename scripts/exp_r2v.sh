#!/bin/bash
# round-2 experiment V: the two-star / 384-thread build (v17) against v16, Philox rounds forced to one IMAD.WIDE, window sweep
cd "$(dirname "$0")/.."
run() { label=$1; shift; timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-trials 1e6 "$@" 2>&1 | tail -1 | python scripts/bench_value.py "$label"; }
cp startrackermpm_b200/libstmpm.so /tmp/libstmpm_keep.so
use() { cp variants/libstmpm_$1.so startrackermpm_b200/libstmpm.so; touch startrackermpm_b200/libstmpm.so; }
for v in v16 v17 v17pw v16 v17 v17pw; do use $v; run "$v default"; done
for v in v17 v17pw; do use $v; run "$v all_visible" --workload all_visible; done
use v17
for w in 192 384 512; do run "v17 win=$w" --window $w; done
echo "v17 mode matrix:"; timeout 600 python scripts/mode_matrix.py 2>&1 | tail -1
echo "v17 gpu tests:"; timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
use v17pw
echo "v17pw parity tests:"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
cp /tmp/libstmpm_keep.so startrackermpm_b200/libstmpm.so; touch startrackermpm_b200/libstmpm.so
