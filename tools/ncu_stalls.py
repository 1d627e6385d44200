#!/usr/bin/env python3
"""Warp stall reasons of a kernel summed over all its instructions (source page of an .ncu-rep).
  python tools/ncu_stalls.py gpurun_out/prof.ncu-rep"""
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--print-source", "sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = next(r for r in rows if r and r[0] == "Address")
cols = [i for i, n in enumerate(hdr) if n.startswith("stall_") and "Not Issued" not in n]
tot = {hdr[i]: 0 for i in cols}
for r in rows:
    if len(r) == len(hdr) and r[0].startswith("0x"):
        for i in cols:
            if r[i].isdigit():
                tot[hdr[i]] += int(r[i])
s = sum(tot.values())
print(f"warp stall samples {s}")
for k, v in sorted(tot.items(), key=lambda t: -t[1]):
    if v:
        print(f"  {k[6:]:20s} {100.0 * v / s:5.1f}%")
