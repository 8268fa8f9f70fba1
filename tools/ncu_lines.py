"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line."""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.01
hdr = None
cur_file = ""
agg = defaultdict(lambda: [0, 0, 0, 0, ""])   # samples, instr, wavefronts, ideal
tot_s = tot_i = 0
for r in rows:
    if len(r) == 2:
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = {n: i for i, n in enumerate(r)}
        # second 'Source' column is SASS; first is CUDA
        continue
    if hdr is None:
        continue
    line, src = r[0], r[1]
    if not line.strip():
        continue          # per-SASS rows repeat what the per-line rows aggregate
    try:
        s = int(r[hdr["# Samples"]] or 0)
        ins = int(r[hdr["Instructions Executed"]] or 0)
        wf = int(r[hdr["L1 Wavefronts Shared"]] or 0)
        wfi = int(r[hdr["L1 Wavefronts Shared Ideal"]] or 0)
    except (ValueError, KeyError):
        continue
    k = (cur_file, line)
    a = agg[k]
    a[0] += s; a[1] += ins; a[2] += wf; a[3] += wfi
    if src.strip():
        a[4] = src.strip()
    tot_s += s; tot_i += ins
print(f"total samples {tot_s}  total warp-instructions {tot_i}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    if a[0] >= thr * tot_s or a[1] >= thr * tot_i:
        print(f"{k[0]}:{k[1]:>4}  samp {100*a[0]/max(tot_s,1):5.1f}%  inst {100*a[1]/max(tot_i,1):5.1f}%  smem_wf {a[2]:>12} ideal {a[3]:>12}  | {a[4][:100]}")
