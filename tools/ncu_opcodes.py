#!/usr/bin/env python3
"""Executed-SASS opcode histogram of an .ncu-rep: python tools/ncu_opcodes.py rep [top_n]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[1]
ie, src, te = h.index("Instructions Executed"), h.index("Source"), h.index("Thread Instructions Executed")
c, t = collections.Counter(), collections.Counter()
for r in rows[2:]:
    if len(r) <= ie or not r[ie].isdigit():
        continue
    s = r[src].strip()
    if s.startswith("@"):
        s = s.split(None, 1)[1]
    op = s.split()[0].split(".")[0]
    c[op] += int(r[ie])
    t[op] += int(r[te])
tot = sum(c.values())
print("total warp instructions", tot)
for op, n in c.most_common(top):
    print(f"{op:12s} {n:10d} {100 * n / tot:5.1f}%  thr {t[op] / max(n, 1):.1f}")
