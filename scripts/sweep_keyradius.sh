#!/bin/bash
# device-resident throughput against the radius (deg) of the work-key disc (0 = candidate-list visits, the v11 key)
for r in ${RADII:-0 12 10.5 9.4 8.5}; do
  timeout 200 python bench.py --key-radius-deg $r --steps 3 --warmup 3 --no-cpu-baseline --e2e-trials 1e6 2>&1 | tail -1 | python scripts/bench_value.py "key_radius=$r"
done
