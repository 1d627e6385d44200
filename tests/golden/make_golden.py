#!/usr/bin/env python3
"""Generate golden CSVs by running the UNMODIFIED reference (/root/reference)
in the build container.  /root/reference does not exist on the GPU box, so the
outputs are committed under tests/golden/cases/ (gzip) together with a manifest
recording the environment they were produced in.

The runner only wraps the reference from the outside:
  * time.sleep is stubbed (the reference sleeps 5 s per chromosome, CROPSR.py:478),
  * numpy's legacy global RNG is seeded so crispr_id is reproducible (:316-318),
  * OPENBLAS_NUM_THREADS is pinned per case (summation order),
  * cases with a `chunk` run the reference's source with ONE literal changed in memory: the
    1000000 of the row-chunk plan (CROPSR.py:451, both occurrences on that line) becomes the
    case's value, so that the plan's quirks (misplaced last slice, dropped exact multiple,
    ids[start-k-1]) are walked on small fixtures.  The file on disk is never touched.

usage: python tests/golden/make_golden.py [case ...]
"""
import gzip
import hashlib
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
FIX = os.path.join(HERE, "fixtures")
OUT = os.path.join(HERE, "cases")
REF = "/root/reference"

RUNNER = r"""
import sys, runpy, time, os
import numpy as np
time.sleep = lambda s: None
np.random.seed(int(os.environ['GOLDEN_SEED']))
sys.path.insert(0, %r)
path = %r
chunk = os.environ.get('GOLDEN_CHUNK')
if chunk is None:
    runpy.run_path(path, run_name='__main__')
else:
    src = open(path).read()
    assert src.count('1000000') == 2
    exec(compile(src.replace('1000000', str(int(chunk))), path, 'exec'), {'__name__': '__main__', '__file__': path})
""" % (REF, os.path.join(REF, "CROPSR.py"))

#        name                     fasta                    -l  seed threads
CASES = [
    ("sample",                "sample_genome.fa",          20, 11, 1),
    ("sample_t6",             "sample_genome.fa",          20, 11, 6),
    ("multi3",                "multi3.fa",                 20, 12, 1),
    ("multi3_l18",            "multi3.fa",                 18, 13, 1),
    ("multi3_l23",            "multi3.fa",                 23, 14, 1),
    ("clean3",                "clean3.fa",                 20, 15, 1),
    ("clean3_trailing_nl",    "clean3_trailing_nl.fa",     20, 16, 1),
    ("edge_clean",            "edge_clean.fa",             20, 17, 1),
    ("edge_fmt",              "edge_fmt.fa",               20, 18, 1),
    ("single_candidate",      "single_candidate.fa",       20, 19, 1),
    ("ws_header",             "ws_header.fa",              20, 20, 1),
    ("dup_keys",              "dup_keys.fa",               20, 21, 1),
    ("empty_records",         "empty_records.fa",          20, 22, 1),
    ("mid50k",                "mid50k.fa",                 20, 23, 1),
    ("mid50k_t5",             "mid50k.fa",                 20, 23, 5),
    # the 1,000,000-row chunk plan shrunk (see the module docstring): name, fasta, -l, seed, threads, chunk
    ("mid50k_c1000",          "mid50k.fa",                 20, 24, 1, 1000),     # q = 4, r = 671: last slice at r*q
    ("mid50k_c1557",          "mid50k.fa",                 20, 25, 1, 1557),     # 4671 = 3 * 1557: final slice dropped
    ("multi3_c20",            "multi3.fa",                 20, 26, 1, 20),       # cumulative lists 50 / 79 / 101
    ("sample_c5000_t4",       "sample_genome.fa",          20, 27, 4, 5000),     # multi-threaded gemv per 5000-row slice
]


def run_case(name, fasta, guide_len, seed, threads, chunk=None):
    with tempfile.TemporaryDirectory() as wd:
        env = dict(os.environ, GOLDEN_SEED=str(seed), OPENBLAS_NUM_THREADS=str(threads))
        if chunk is not None:
            env["GOLDEN_CHUNK"] = str(chunk)
        argv = [sys.executable, "-c", RUNNER, "-f", os.path.join(FIX, fasta),
                "-g", os.path.join(FIX, "sample_genome.gff"), "-o", os.path.join(wd, "out.csv"),
                "-l", str(guide_len), "--cas9"]
        p = subprocess.run(argv, cwd=wd, env=env, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError(f"{name}: reference failed\n{p.stderr}")
        with open(os.path.join(wd, "out.csv"), "rb") as f:
            data = f.read()
        time_txt = open(os.path.join(wd, "time.txt")).read()
    with gzip.GzipFile(os.path.join(OUT, name + ".csv.gz"), "wb", mtime=0) as f:
        f.write(data)
    with open(os.path.join(OUT, name + ".stdout"), "w") as f:
        f.write(p.stdout)
    return {"fasta": fasta, "guide_len": guide_len, "seed": seed, "blas_threads": threads, "chunk": chunk,
            "csv_sha256": hashlib.sha256(data).hexdigest(), "rows": data.count(b"\r\n") - 1,
            "time_txt_records": time_txt.count("Total runtime of the program is ")}


def main():
    import numpy
    os.makedirs(OUT, exist_ok=True)
    only = set(sys.argv[1:])
    mpath = os.path.join(OUT, "manifest.json")
    manifest = json.load(open(mpath)) if os.path.exists(mpath) else {"cases": {}}
    for case in CASES:
        if only and case[0] not in only:
            continue
        manifest["cases"][case[0]] = run_case(*case)
        print(case[0], manifest["cases"][case[0]]["rows"], "rows")
    manifest["environment"] = {
        "numpy": numpy.__version__,
        "python": sys.version.split()[0],
        "cpu": next((l.split(":", 1)[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")), "?"),
        "note": "scipy-openblas 0.3.30 DYNAMIC_ARCH; AVX-512 host (np.exp SIMD path)",
    }
    json.dump(manifest, open(mpath, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
