"""Join an `ncu --page source --csv --print-source sass` export with `nvdisasm --print-line-info` of the same kernel and
aggregate warp-instructions / stall samples by CUDA source line.
Usage: python scripts/ncu_by_line.py sass.csv disasm.txt [n_trials] [top]"""
import csv, re, sys, collections, os
sass_csv, dis = sys.argv[1], sys.argv[2]
n_trials = float(sys.argv[3]) if len(sys.argv) > 3 else None
top = int(sys.argv[4]) if len(sys.argv) > 4 else 60
line_of = {}
cur = ("?", 0)
for l in open(dis):
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", l)
    if m: line_of[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(sass_csv)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
ix = {h: i for i, h in enumerate(rows[hi])}
base = None
agg = collections.defaultdict(lambda: [0, 0, 0])
stall_cols = [h for h in rows[hi] if h.startswith("stall_") and "Not Issued" not in h]
stall_by = collections.defaultdict(lambda: collections.Counter())
tot = [0, 0, 0]
for r in rows[hi + 1:]:
    if len(r) < len(ix): continue
    a = int(r[0], 16)
    if base is None: base = a
    key = line_of.get(a - base, ("?", 0))
    try:
        inst = int(float(r[ix["Instructions Executed"]] or 0)); thr = int(float(r[ix["Thread Instructions Executed"]] or 0))
        samp = int(float(r[ix["# Samples"]] or 0))
    except ValueError:
        continue
    agg[key][0] += inst; agg[key][1] += thr; agg[key][2] += samp
    for h in stall_cols:
        v = r[ix[h]]
        if v and v != "0": stall_by[key][h[6:]] += int(float(v))
    tot[0] += inst; tot[1] += thr; tot[2] += samp
print(f"total warp-inst {tot[0]:.3e} samples {tot[2]}")
print(f"{'file:line':28s} {'warp-inst%':>10s} {'lanes':>6s} {'samples%':>8s}  top stalls")
for key, (i, t, s) in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
    st = ", ".join(f"{k}={100 * v / max(s, 1):.0f}%" for k, v in stall_by[key].most_common(3))
    print(f"{key[0] + ':' + str(key[1]):28s} {100 * i / tot[0]:10.2f} {t / max(i, 1):6.1f} {100 * s / tot[2]:8.2f}  {st}")
