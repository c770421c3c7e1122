#!/usr/bin/env python
"""Dump a SASS range of an .ncu-rep with per-instruction stall samples converted to cycles.
usage: sass_regions.py rep lo hi nwarps_executing [min_cycles] [kernel-index]
cycles = samples / (samples-per-cycle), calibrated on the 'selected' samples of the range
(one issue cycle each)."""
import csv, io, subprocess, sys
rep, lo, hi = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
thr = float(sys.argv[4]) if len(sys.argv) > 4 else 8
kidx = int(sys.argv[5]) if len(sys.argv) > 5 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
sec = rows[starts[kidx]:starts[kidx + 1]]
h = sec[1]; data = [r for r in sec[2:] if len(r) == len(h)]
iS = h.index("# Samples"); iE = h.index("Instructions Executed"); iSrc = h.index("Source")
st = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
iSel = h.index("stall_selected")
sel = data[lo:hi]
# calibrate: median 'selected' samples of instructions executed the modal number of times
import statistics
ex = statistics.mode(float(r[iE]) for r in sel)
per_cycle = statistics.median(float(r[iSel]) for r in sel if float(r[iE]) == ex and float(r[iSel]) > 0)
print(f"modal executions {ex:.0f}, samples per issue cycle {per_cycle:.1f}")
cum = 0.0
for k, r in enumerate(sel):
    s = float(r[iS]); cum += s
    if s / per_cycle >= thr or 'BAR' in r[iSrc] or 'SYNCS' in r[iSrc]:
        mix = sorted(((float(r[i]), h[i][6:]) for i in st), reverse=True)[:2]
        print(f"{lo+k:5d} {r[iSrc].strip()[:58]:58s} x{float(r[iE])/ex:4.2f} cyc={s/per_cycle:6.0f} cum={cum/per_cycle:6.0f} "
              + " ".join(f"{n}={v/per_cycle:.0f}" for v, n in mix if v > 0))
print(f"total {cum/per_cycle:.0f} cycles")
