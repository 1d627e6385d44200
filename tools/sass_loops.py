#!/usr/bin/env python3
"""Static size of the loops of k_scan_score<true> in the built library (no GPU needed).

  python tools/sass_loops.py [lib.so]

Lists every backward branch (loop) of the kernel with its instruction count and opcode mix;
the two loops with DADDs are the per-candidate bodies ('+' and '-' strand).
"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "cropsr_b200/libcropsr_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
fn = None
ins = []
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        continue
    if fn and "k_scan_scoreILb1" in fn:
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
print(f"{len(ins)} instructions")
addr_index = {a: i for i, (a, _) in enumerate(ins)}
for i, (a, s) in enumerate(ins):
    m = re.search(r"BRA(?:\.\S+)?\s+(?:\S+,\s*)?0x([0-9a-f]+)", s)
    if not m:
        continue
    tgt = int(m.group(1), 16)
    if tgt <= a and tgt in addr_index:
        j = addr_index[tgt]
        body = [x[1] for x in ins[j:i + 1]]
        ops = collections.Counter(re.sub(r"^@!?U?P\d\s+", "", b).split()[0].split(".")[0] for b in body)
        if len(body) > 40:
            if ops.get("DADD") and len(body) < 250:      # a candidate body: instructions per issue pipe
                alu = sum(ops.get(k, 0) for k in ("LOP3", "SHF", "IADD3", "LEA", "ISETP", "VIADD", "VIADDMNMX", "SEL", "PRMT", "PLOP3", "IABS", "BREV"))
                fma = sum(v for k, v in ops.items() if k.startswith("IMAD") or k in ("FFMA", "FMUL", "FADD"))
                print(f"  body: ALU pipe {alu}  FMA pipe {fma}  FP64 {ops.get('DADD', 0) + ops.get('DFMA', 0) + ops.get('DMUL', 0)}  "
                      f"LDS {ops.get('LDS', 0)}  STG {ops.get('STG', 0)}  total {len(body)}")
            print(f"loop {tgt:#x}..{a:#x}: {len(body)} instr  " + " ".join(f"{k}={v}" for k, v in ops.most_common(14)))
