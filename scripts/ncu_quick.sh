#!/bin/bash
# scripts/ncu_quick.sh NAME...: a handful of ncu counters of the trial kernel for each variants/libstmpm_NAME.so (after a plain run)
cd "$(dirname "$0")/.."
cp startrackermpm_b200/libstmpm.so /tmp/libstmpm_keep.so
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_op_read.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,gpu__time_duration.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,sm__cycles_elapsed.avg
CMD="python bench.py --steps 2 --warmup 3 --trials 2e7 --e2e-trials 1e6 --no-cpu-baseline"
for v in "$@"; do
  cp variants/libstmpm_$v.so startrackermpm_b200/libstmpm.so
  $CMD > /dev/null 2>&1 && ncu --metrics $M --clock-control none -k regex:trials_kernel -s 3 -c 1 --csv --log-file gpurun_out/ncuq_$v.csv $CMD > /dev/null 2>&1
  python - "$v" <<'PY'
import csv,sys
v=sys.argv[1]
rows=[r for r in csv.reader(open(f'gpurun_out/ncuq_{v}.csv')) if len(r)>14 and r[0]=='0']
d={r[12]:r[14] for r in rows}
short=lambda k:k.replace('smsp__average_warps_issue_stalled_','st_').replace('_per_issue_active.ratio','')
print(v,' '.join(f"{short(k)}={x}" for k,x in d.items()))
PY
done
cp /tmp/libstmpm_keep.so startrackermpm_b200/libstmpm.so
