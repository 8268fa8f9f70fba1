"""Opcode histogram (weighted by executed warp-instructions and by stall samples) of an
`ncu --page source --csv --print-source sass` dump, plus the SASS lines of the hottest code."""
import csv
import re
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hdr = None
ops = defaultdict(lambda: [0, 0])
lines = []
for r in rows:
    if "Instructions Executed" in r and "Source" in r:
        hdr = {n: i for i, n in enumerate(r)}
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    try:
        ins = int(r[hdr["Instructions Executed"]] or 0)
        smp = int(r[hdr["# Samples"]] or 0)
    except ValueError:
        continue
    src = r[hdr["Source"]].strip()
    m = re.match(r"(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", src)
    if not m:
        continue
    op = m.group(1)
    ops[op][0] += ins
    ops[op][1] += smp
    lines.append((ins, smp, src))
ti = sum(v[0] for v in ops.values()) or 1
ts = sum(v[1] for v in ops.values()) or 1
print(f"total warp-instructions {ti}  samples {ts}")
for op, (i, s) in sorted(ops.items(), key=lambda kv: -kv[1][0])[:40]:
    print(f"{op:12s} inst {100 * i / ti:5.1f}%  samp {100 * s / ts:5.1f}%")
if top:
    print("---- SASS lines by samples")
    for ins, smp, src in sorted(lines, key=lambda t: -t[1])[:top]:
        print(f"samp {100 * smp / ts:5.2f}%  inst {100 * ins / ti:5.2f}%  {src[:110]}")
