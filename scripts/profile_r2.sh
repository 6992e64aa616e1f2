#!/bin/bash
# scripts/profile_r2.sh — round-2 ncu evidence (run under gpurun, one GPU).  Every ncu run follows a plain run of the SAME
# command that exited 0.  Outputs under gpurun_out/: r2_launches.csv, r2_{fp64,fp32,allvis,rng}.ncu-rep + plain logs.
cd "$(dirname "$0")/.."
set -o pipefail
B="python bench.py --steps 2 --warmup 3 --trials 2e7 --e2e-trials 1e6 --no-cpu-baseline"
NCU="ncu --clock-control none"
# A/B on this box first (plain timings): committed build vs the no-clamp-only variant of experiment A
run() { label=$1; shift; timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-trials 1e6 "$@" 2>&1 | tail -1 | python scripts/bench_value.py "$label"; }
run "committed build, default"
run "committed build, all_visible" --workload all_visible
# 1. launch list of a short default bench run (the kernel's SHARE of the step)
$B > gpurun_out/r2_plain_default.log 2>&1 && \
  $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/r2_launches.csv $B > gpurun_out/r2_ncu_launches.log 2>&1
# 2. full capture of the pre-sampled fp64 kernel (bench launches it first in each step: skip the 3 warm-ups)
$B > /dev/null 2>&1 && \
  $NCU --set full --import-source on --kernel-name-base mangled -k regex:trials_kernelILi0ELi64ELi0 -s 3 -c 1 -f -o gpurun_out/r2_fp64 $B > gpurun_out/r2_ncu_fp64.log 2>&1
# 3. the fp32 instantiation (bench's fp32 block)
$B > /dev/null 2>&1 && \
  $NCU --set full --import-source on --kernel-name-base mangled -k regex:trials_kernelILi0ELi32ELi0 -s 1 -c 1 -f -o gpurun_out/r2_fp32 $B > gpurun_out/r2_ncu_fp32.log 2>&1
# 4. star-rich fields (MAX_MAGNITUDE = 20)
$B --workload all_visible > /dev/null 2>&1 && \
  $NCU --set full --import-source on --kernel-name-base mangled -k regex:trials_kernelILi0ELi64ELi0 -s 3 -c 1 -f -o gpurun_out/r2_allvis $B --workload all_visible > gpurun_out/r2_ncu_allvis.log 2>&1
# 5. native-RNG instantiation (configs 3-5)
python scripts/profile_rng.py > /dev/null 2>&1 && \
  $NCU --set full --import-source on --kernel-name-base mangled -k regex:trials_kernelILi1ELi64ELi0 -s 2 -c 1 -f -o gpurun_out/r2_rng python scripts/profile_rng.py > gpurun_out/r2_ncu_rng.log 2>&1
ls -la gpurun_out/*.ncu-rep
