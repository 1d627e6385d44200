#!/usr/bin/env python3
"""Scan time of the library selected by CROPSR_B200_LIB (kernel experiments).
usage: CROPSR_B200_LIB=tools/variants/x.so python tools/variant_scan.py [workload] [n] [flags]
Prints the median / min CUDA-event time of n scans, L2 flushed between scans, and a digest
of the candidate streams (so that a timing-only variant that breaks parity is visible)."""
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from cropsr_b200 import engine

workload = sys.argv[1] if len(sys.argv) > 1 else "arabidopsis"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 12
flags = int(sys.argv[3]) if len(sys.argv) > 3 else 0
engine.init(0)
g = engine.Genome()
for t in bench.synth_tokens(workload):
    g.add_token(t)
g.commit()
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
ms = []
digest = None
for i in range(n + 3):
    flush.zero_()
    torch.cuda.synchronize()
    r = g.scan(20, flags)
    if i >= 3:
        ms.append(r.scan_ms())
    if i == n + 2:
        h = hashlib.sha256()
        for strand in "+-":
            f = r.fetch(strand)
            for k in ("pos", "packed", "x"):
                if k in f and f[k] is not None:
                    h.update(np.ascontiguousarray(f[k]).tobytes())
        digest = h.hexdigest()[:16]
    r.free()
ms = np.array(ms) * 1e3
print(f"{os.environ.get('CROPSR_B200_LIB', 'default'):<40} {workload} median {np.median(ms):7.2f} us  min {ms.min():7.2f} us  digest {digest}  pack {g.timing()['pack_ms'] * 1e3:.1f} us")
