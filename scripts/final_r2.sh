#!/bin/bash
# scripts/final_r2.sh — the evidence run of the committed build (under gpurun, one GPU): GPU tests, smoke, bench lines, mode matrix,
# then the ncu captures of scripts/profile_r2.sh.  Outputs under gpurun_out/.
cd "$(dirname "$0")/.."
timeout 1200 python -m pytest tests -m gpu -q -rs > gpurun_out/r2_gputest.log 2>&1; tail -2 gpurun_out/r2_gputest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -1 gpurun_out/r2_smoke.log
timeout 600 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -c 600 gpurun_out/r2_bench_n1.json
timeout 300 python bench.py --workload all_visible --no-cpu-baseline > gpurun_out/r2_bench_allvis.json 2>/dev/null
timeout 300 python scripts/mode_matrix.py 2>/dev/null | tail -1 > gpurun_out/r2_mode_matrix.json; cat gpurun_out/r2_mode_matrix.json
bash scripts/profile_r2.sh > gpurun_out/r2_profile.log 2>&1; tail -8 gpurun_out/r2_profile.log
