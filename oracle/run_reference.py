#!/usr/bin/env python3
"""Run the UNMODIFIED reference CLI (baseline/_ref/CROPSR.py, see install_reference.py) from the
outside, the way SURVEY.md Appendix B does: a fresh interpreter, `time.sleep` stubbed (the
reference sleeps 5 s per chromosome, CROPSR.py:478), numpy's legacy RNG optionally seeded (ids,
CROPSR.py:316-318), cwd = a scratch directory (it writes time.txt there, CROPSR.py:371).

TEST / BENCH INFRASTRUCTURE ONLY (bench.py's --impl reference and cpu_baseline legs, the golden
generators).  Nothing under cropsr_b200/ imports this.

    run(fasta, gff, out_csv, workdir, guide_len=20, seed=None, threads=None) -> dict
        wall_s        wall clock of the whole interpreter run (imports included)
        program_s     what the reference itself reports: the last "Total runtime" of time.txt
        blas_threads  OpenBLAS threads inside np.matmul (the only multi-threaded part)
"""
import json
import os
import re
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "baseline", "_ref")

_RUNNER = r"""
import sys, runpy, time, os, json
import numpy as np
time.sleep = lambda s: None
seed = os.environ.get('CROPSR_REF_SEED')
if seed is not None:
    np.random.seed(int(seed))
ref = os.environ['CROPSR_REF_DIR']
sys.path.insert(0, ref)
threads = None
try:
    from threadpoolctl import threadpool_info
    threads = max([p.get('num_threads', 1) for p in threadpool_info() if p.get('user_api') == 'blas'] or [1])
except Exception:
    pass
t0 = time.perf_counter()
runpy.run_path(os.path.join(ref, 'CROPSR.py'), run_name='__main__')
sys.stderr.write('\n@@REF ' + json.dumps({'main_s': time.perf_counter() - t0, 'blas_threads': threads}) + '\n')
"""


def available():
    return os.path.exists(os.path.join(REF_DIR, "CROPSR.py")) and os.path.exists(os.path.join(REF_DIR, "cropsr_functions.py"))


def run(fasta, gff, out_csv, workdir, guide_len=20, seed=None, threads=None, timeout=None):
    env = dict(os.environ, CROPSR_REF_DIR=REF_DIR)
    if seed is not None:
        env["CROPSR_REF_SEED"] = str(seed)
    if threads is not None:
        env["OPENBLAS_NUM_THREADS"] = str(threads)
    argv = [sys.executable, "-c", _RUNNER, "-f", fasta, "-g", gff, "-o", out_csv, "-l", str(guide_len), "--cas9"]
    t0 = time.perf_counter()
    p = subprocess.run(argv, cwd=workdir, env=env, capture_output=True, text=True, timeout=timeout)
    wall = time.perf_counter() - t0
    if p.returncode != 0:
        raise RuntimeError("reference CLI failed:\n" + p.stderr[-2000:])
    info = {}
    for line in p.stderr.splitlines():
        if line.startswith("@@REF "):
            info = json.loads(line[6:])
    with open(os.path.join(workdir, "time.txt")) as f:
        stamps = re.findall(r"Total runtime of the program is ([0-9.eE+-]+?)(?=Total|$)", f.read())
    return {"wall_s": wall, "main_s": info.get("main_s"), "program_s": float(stamps[-1]) if stamps else None,
            "blas_threads": info.get("blas_threads"), "stdout": p.stdout}


if __name__ == "__main__":
    import tempfile
    fasta, gff = sys.argv[1], sys.argv[2]
    with tempfile.TemporaryDirectory() as wd:
        r = run(os.path.abspath(fasta), os.path.abspath(gff), os.path.join(wd, "out.csv"), wd)
        r.pop("stdout")
        print(json.dumps(r))
