#!/usr/bin/env python3
"""Wall time of the drop-in CLI path (pipeline.run_cas9) on a synthetic multi-line FASTA,
with and without the device-side ingest.  usage: python tools/cli_e2e.py [Mbp per chromosome] [chromosomes]"""
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from cropsr_b200 import engine, pipeline

mbp = float(sys.argv[1]) if len(sys.argv) > 1 else 10
nchr = int(sys.argv[2]) if len(sys.argv) > 2 else 2
engine.init(0)
tmp = tempfile.mkdtemp()
fa, gff = os.path.join(tmp, "g.fa"), os.path.join(tmp, "g.gff")
rng = np.random.default_rng(1)
with open(fa, "wb") as f:
    for k in range(nchr):
        n = int(mbp * 1e6) // 80 * 80
        s = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, n)].reshape(-1, 80)
        f.write(f">chr{k + 1}\n".encode())
        f.write(np.concatenate([s, np.full((len(s), 1), 10, np.uint8)], axis=1).tobytes())
open(gff, "w").write("##gff-version 3\nchr1\tsyn\tgene\t100\t900\t.\t+\t.\tID=g1\n")
for dev in (True, False, True):
    np.random.seed(1)
    t0 = time.time()
    stats = pipeline.run_cas9(fa, gff, os.path.join(tmp, "out.csv"), 20, False, 1, os.path.join(tmp, "time.txt"),
                              out=lambda *a: None, device_ingest=dev)
    dt = time.time() - t0
    print(f"device_ingest={dev}: {dt:.2f} s wall, {stats['candidates']} candidates, {stats['rows']} rows, "
          f"csv {os.path.getsize(os.path.join(tmp, 'out.csv')) / 1e6:.0f} MB, scan {stats['scan_ms']:.3f} ms, "
          f"h2d {stats['h2d_ms']:.2f} ms, pack {stats['pack_ms']:.3f} ms")
