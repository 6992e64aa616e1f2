"""Aggregate an `ncu --page source --csv` export by SASS opcode: warp-instructions executed, average active threads,
stall samples.  Usage: python scripts/ncu_opmix.py src.csv [n_trials]"""
import csv, sys, collections, re
path = sys.argv[1]
n_trials = float(sys.argv[2]) if len(sys.argv) > 2 else None
rows = list(csv.reader(open(path)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
ix = {h: i for i, h in enumerate(hdr)}
ops = collections.defaultdict(lambda: [0, 0, 0])
tot_inst = tot_thr = tot_samp = 0
for r in rows[hdr_i + 1:]:
    if len(r) < len(hdr): continue
    src = r[ix["Source"]].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    if not m: continue
    op = m.group(2)
    base = ".".join(op.split(".")[:2]) if op.split(".")[0] in ("MUFU", "F2F", "I2F", "F2I", "LDS", "LDG", "STS", "ATOMS", "DSETP", "FSETP") else op.split(".")[0]
    try:
        inst = int(float(r[ix["Instructions Executed"]] or 0)); thr = int(float(r[ix["Thread Instructions Executed"]] or 0))
        samp = int(float(r[ix["# Samples"]] or 0))
    except ValueError:
        continue
    ops[base][0] += inst; ops[base][1] += thr; ops[base][2] += samp
    tot_inst += inst; tot_thr += thr; tot_samp += samp
print(f"total warp-inst {tot_inst:.3e}  thread-inst {tot_thr:.3e}  avg active lanes {tot_thr / max(tot_inst, 1):.1f}  samples {tot_samp}")
if n_trials: print(f"per trial: {tot_thr / n_trials:.0f} thread-inst, {tot_inst * 32 / n_trials:.0f} issue-slots x32")
print(f"{'op':14s} {'warp-inst':>12s} {'%':>6s} {'lanes':>6s} {'samples%':>8s}" + ("  thr/trial" if n_trials else ""))
for op, (i, t, s) in sorted(ops.items(), key=lambda kv: -kv[1][0])[:45]:
    print(f"{op:14s} {i:12d} {100 * i / tot_inst:6.2f} {t / max(i, 1):6.1f} {100 * s / max(tot_samp, 1):8.2f}" + (f"  {t / n_trials:8.1f}" if n_trials else ""))
