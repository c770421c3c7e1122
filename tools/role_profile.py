#!/usr/bin/env python
"""Role-level view of a scan_fast ncu report: splits the SASS of the kernel into segments at
named-barrier instructions, and prints for each segment its share of the warp-stall samples,
executed instructions per time step and the stall mix.  The barrier that a segment starts with
tells which role it belongs to (bar ids: 1,2 OPS  3,4 PART  5 CH  6,7 FULL  8,9 EMPTY).

usage: tools/role_profile.py rep.ncu-rep steps_per_sm [kernel-index]
"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
steps = float(sys.argv[2])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# one section per profiled kernel: "Kernel Name" row, header row, instruction rows
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
kidx = int(sys.argv[3]) if len(sys.argv) > 3 else 0
sec = rows[starts[kidx]:starts[kidx + 1]]
print("kernel:", sec[0][1][:80])
h = sec[1]
data = [r for r in sec[2:] if len(r) == len(h)]
iS, iE, iSrc = h.index("# Samples"), h.index("Instructions Executed"), h.index("Source")
stalls = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
tot = sum(float(r[iS]) for r in data)
nsm = 148.0


def is_fp64(src):
    op = src.strip().split()
    op = [x for x in op if not x.startswith("@")]
    return bool(op) and op[0].split(".")[0] in ("DFMA", "DADD", "DMUL", "DSETP")


seg_start = 0
print(f"total samples {tot:.0f}")
print(f"{'range':>11s} {'samp%':>6s} {'instr/step':>10s} {'fp64/step':>9s}  first-instr / stall mix")
for k, r in enumerate(data + [None]):
    last = r is None
    src = "" if last else r[iSrc]
    if last or "BAR." in src or "EXIT" in src:
        sel = data[seg_start:k + (0 if last else 1)]
        smp = sum(float(x[iS]) for x in sel)
        ex = sum(float(x[iE]) for x in sel) / (nsm * steps)
        fp = sum(float(x[iE]) for x in sel if is_fp64(x[iSrc])) / (nsm * steps)
        if smp / max(tot, 1) > 0.002:
            agg = {}
            for x in sel:
                for i in stalls:
                    agg[h[i]] = agg.get(h[i], 0) + float(x[i])
            mix = ", ".join(f"{n[6:]}={100 * v / max(smp, 1):.0f}" for n, v in
                            sorted(agg.items(), key=lambda kv: -kv[1])[:5])
            print(f"{seg_start:5d}-{k:5d} {100 * smp / tot:6.2f} {ex:10.1f} {fp:9.1f}  "
                  f"[{'end' if last else src.strip()[:40]}] {mix}")
        seg_start = k + 1
