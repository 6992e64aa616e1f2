#!/bin/bash
# rebuild with different threads-per-CTA (register cap follows from __launch_bounds__) and time the bench
for t in ${TPBS:-512 576 640}; do
  sed -i "s/^constexpr int TPB = [0-9]*;/constexpr int TPB = $t;/" startrackermpm_b200/csrc/stmpm_device.cuh
  python -m startrackermpm_b200.build 2>&1 | grep -E "error" 
  timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-trials 1e6 2>&1 | tail -1 | python scripts/bench_value.py "tpb=$t"
done
sed -i "s/^constexpr int TPB = [0-9]*;/constexpr int TPB = 512;/" startrackermpm_b200/csrc/stmpm_device.cuh
