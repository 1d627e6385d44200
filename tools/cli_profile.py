#!/usr/bin/env python3
"""cProfile of the drop-in CLI path on a synthetic FASTA: python tools/cli_profile.py [Mbp per chromosome] [chromosomes]"""
import cProfile
import os
import pstats
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
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
run = lambda: pipeline.run_cas9(fa, gff, os.path.join(tmp, "out.csv"), 20, False, 1, os.path.join(tmp, "time.txt"),
                                out=lambda *a: None)
np.random.seed(1)
run()                                   # warm
np.random.seed(1)
pr = cProfile.Profile()
t0 = time.time()
pr.enable()
stats = run()
pr.disable()
print(f"{time.time() - t0:.2f} s wall, {stats['rows']} rows")
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
