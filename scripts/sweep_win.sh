#!/bin/bash
# device-resident throughput against the CTA sort-window size (trials)
for w in ${WINS:-2048 1024 512 256 128}; do
  timeout 200 python bench.py --window $w --steps 3 --warmup 3 --no-cpu-baseline --e2e-trials 1e6 2>&1 | tail -1 | python scripts/bench_value.py "win=$w"
done
