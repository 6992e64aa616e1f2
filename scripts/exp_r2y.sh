#!/bin/bash
# round-2 experiment Y: unconditional clamped record prefetch in the two-star loop (y1), + two-way unrolled (y2)
cd "$(dirname "$0")/.."
run() { label=$1; shift; timeout 200 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-trials 1e6 "$@" 2>&1 | tail -1 | python scripts/bench_value.py "$label"; }
cp startrackermpm_b200/libstmpm.so /tmp/libstmpm_keep.so
use() { cp variants/libstmpm_$1.so startrackermpm_b200/libstmpm.so; touch startrackermpm_b200/libstmpm.so; }
for v in ${VARIANTS:-y0 y1 y2 y0 y1}; do use $v; run "$v default"; done
use y1; echo "y1 parity + scheduler stress:"; timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -m gpu -q -x -k "not parity_at_scale" 2>&1 | tail -2
cp /tmp/libstmpm_keep.so startrackermpm_b200/libstmpm.so; touch startrackermpm_b200/libstmpm.so
