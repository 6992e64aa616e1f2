#!/bin/bash
# scripts/make_profile_summaries.sh — turns gpurun_out/r2_{fp64,fp32,allvis,rng}.ncu-rep (scripts/profile_r2.sh) into the committed
# text summaries under profiles/ (raw metric summary, SASS opcode mix, stall samples by CUDA source line) and profiles/traffic.json.
# The by-line join needs the .so that was profiled: run right after the capture, before rebuilding.
cd "$(dirname "$0")/.."
set -e
rm -rf /tmp/stmpm_cub && mkdir /tmp/stmpm_cub && (cd /tmp/stmpm_cub && cuobjdump -xelf all "${STMPM_SO:-$OLDPWD/startrackermpm_b200/libstmpm.so}" > /dev/null)
declare -A FN=( [fp64]=_Z13trials_kernelILi0ELi64ELi0EEv5KArgs [fp32]=_Z13trials_kernelILi0ELi32ELi0EEv5KArgs [allvis]=_Z13trials_kernelILi0ELi64ELi0EEv5KArgs [rng]=_Z13trials_kernelILi1ELi64ELi0EEv5KArgs )
for k in fp64 fp32 allvis rng; do
  rep=gpurun_out/r2_$k.ncu-rep
  [ -f $rep ] || continue
  python scripts/ncu_summary.py $rep > profiles/r2_${k}_raw_summary.txt
  ncu -i $rep --page source --csv --print-source sass > /tmp/r2_${k}_sass.csv 2>/dev/null
  python scripts/ncu_opmix.py /tmp/r2_${k}_sass.csv 2e7 > profiles/r2_${k}_opmix.txt
  nvdisasm -g -hex /tmp/stmpm_cub/*.cubin 2>/dev/null | awk -v f="${FN[$k]}" '/^\/\/-+ \.text\./{on=($0 ~ f)} on' > /tmp/dis_$k.txt
  python scripts/ncu_by_line.py /tmp/r2_${k}_sass.csv /tmp/dis_$k.txt 2e7 45 > profiles/r2_${k}_byline.txt
done
cp gpurun_out/r2_launches.csv profiles/r2_launches.csv
python - <<'PY'
import json, re
def metrics(path):
    d = {}
    for line in open(path):
        m = re.match(r"(\S+) \[(.*?)\] = (\S+)", line)
        if m:
            v = float(m.group(3).replace(",", ""))
            u = m.group(2)
            d[m.group(1)] = v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(u, 1)
    return d
m = metrics("profiles/r2_fp64_raw_summary.txt")
n = 2e7
out = {"capture": "profiles/r2_fp64_raw_summary.txt (ncu --set full, trials_kernel<0,64,0>, 2e7 trials per launch, round 2)",
       "dram_bytes_read": m["dram__bytes_read.sum"], "dram_bytes_write": m["dram__bytes_write.sum"], "trials_per_launch": n,
       "dram_bytes_per_trial": (m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"]) / n,
       "pipes": {k: m[f"sm__inst_executed_pipe_{k}.avg.pct_of_peak_sustained_active"] for k in ("fp64", "fma", "alu", "xu", "lsu")},
       "issue_active_pct": m["smsp__issue_active.avg.pct_of_peak_sustained_active"],
       "lanes": m["smsp__thread_inst_executed_per_inst_executed.ratio"]}
json.dump(out, open("profiles/traffic.json", "w"), indent=1)
print(json.dumps(out))
PY
