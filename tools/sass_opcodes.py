#!/usr/bin/env python
"""Opcode histogram of the hand-written kernels from the built library (cuobjdump -sass; no GPU).
usage: tools/sass_opcodes.py [lib] > profiles/r2_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gadfly_b200", "libgadfly_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
fn, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        fn = re.sub(r"gf::\(anonymous namespace\)::", "", fn)
        hist[fn] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and fn:
        op = m.group(1)
        hist[fn][op.split(".")[0]] += 1
        if op.startswith(("SYNCS", "BAR", "USETMAXREG", "DFMA", "LDS", "STS", "SHFL", "LDL", "STL", "MUFU")):
            hist[fn]["  " + op] += 1
print(f"# SASS opcode counts per kernel of {os.path.basename(lib)} (sm_100a; tools/sass_opcodes.py).")
print("# Static counts (code, not executed instructions).  FP64 work is DFMA/DMUL/DADD; SYNCS.* = mbarrier,")
print("# BAR.* = named barriers, USETMAXREG = setmaxnreg register re-balancing, LDL/STL = local-memory spills.")
print("# No tensor-core (HMMA/UTCMMA/tcgen05) and no TMA (UBLKCP) opcodes: the recursion is FP64 rank-1 updates.")
for fn, h in hist.items():
    if not h:
        continue
    total = sum(v for k, v in h.items() if not k.startswith("  "))
    print(f"\n## {fn}\n   {total} instructions")
    main = [(k, v) for k, v in h.items() if not k.startswith("  ")]
    print("   " + ", ".join(f"{k} {v}" for k, v in sorted(main, key=lambda kv: -kv[1])[:18]))
    det = [(k.strip(), v) for k, v in h.items() if k.startswith("  ")]
    print("   detail: " + ", ".join(f"{k} {v}" for k, v in sorted(det, key=lambda kv: -kv[1])[:24]))
