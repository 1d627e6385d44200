#!/usr/bin/env python3
"""Print the SASS of k_scan_score<true> between two addresses: python tools/sass_dump.py 0x60c0 0x6a20 [lib]"""
import re, subprocess, sys
lo, hi = int(sys.argv[1], 16), int(sys.argv[2], 16)
lib = sys.argv[3] if len(sys.argv) > 3 else "cropsr_b200/libcropsr_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
fn = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1); continue
    if fn and "k_scan_scoreILb1" in fn:
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and lo <= int(m.group(1), 16) <= hi:
            print(m.group(1), m.group(2).strip())
