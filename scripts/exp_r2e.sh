#!/bin/bash
# round-2 experiment E: split pipeline per-kernel times (ncu launch list) + end-to-end value
cd "$(dirname "$0")/.."
run() { label=$1; shift; timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-trials 1e6 "$@" 2>&1 | tail -1 | python scripts/bench_value.py "$label"; }
run "single kernel" --pipeline 0
run "split pipeline" --pipeline 1
B="python bench.py --steps 2 --warmup 3 --trials 2e7 --e2e-trials 1e6 --no-cpu-baseline --pipeline 1"
$B > /dev/null 2>&1 && ncu --clock-control none --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum -k regex:split -s 6 -c 6 --csv --log-file gpurun_out/r2_split_launches.csv $B > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_split_launches.csv')) if len(r)>10]
ix={h:i for i,h in enumerate(rows[0])}
for r in rows[1:]:
    print(r[ix['ID']], r[ix['Kernel Name']][:40], r[ix['Metric Name']], r[ix['Metric Value']], r[ix['Metric Unit']])
PY
