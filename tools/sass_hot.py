#!/usr/bin/env python
"""Per-instruction view of an .ncu-rep source page: sample share of instruction ranges and the
hottest instructions inside a range.  usage: sass_hot.py rep [lo hi [top]]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[1]; data = rows[2:]
iS = h.index("# Samples"); iE = h.index("Instructions Executed"); iSrc = h.index("Source")
stalls = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
tot = sum(float(r[iS]) for r in data)
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else len(data)
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
sel = data[lo:hi]
ssum = sum(float(r[iS]) for r in sel)
print(f"range [{lo},{hi}) : {100*ssum/tot:.1f}% of samples; executed warp-instr {sum(float(r[iE]) for r in sel):.3g}")
agg = {}
for r in sel:
    for i in stalls:
        agg[h[i]] = agg.get(h[i], 0) + float(r[i])
print("stall mix:", ", ".join(f"{k[6:]}={100*v/max(ssum,1):.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
order = sorted(range(len(sel)), key=lambda k: -float(sel[k][iS]))[:top]
for k in sorted(order):
    r = sel[k]
    why = max(stalls, key=lambda i: float(r[i]))
    print(f"{lo+k:5d} {100*float(r[iS])/tot:5.2f}% {h[why][6:]:>12s} {r[iSrc].strip()[:80]}")
