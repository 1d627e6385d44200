#!/usr/bin/env python3
"""Golden DIGESTS of runs of the UNMODIFIED reference with more than 1,000,000 candidates, i.e.
through the 1,000,000-row chunk plan of /root/reference/CROPSR.py:451-472 (SURVEY.md 8a row 10:
the last partial slice starts at r*q instead of 1e6*q, ids are indexed ids[start-k-1]) and, with
8 OpenBLAS threads, through the per-thread row classes of a 1,000,000-row np.matmul.

The CSVs are 150-250 MB, so only their sha256, row count and the two normalised digests of
SURVEY.md section 4 are committed (tests/golden/cases/big_manifest.json); the FASTA is
regenerated from the seed by `big_fasta()` below, which the GPU parity test imports.

usage: python tests/golden/make_big_golden.py [case ...]      (needs /root/reference, ~15 GB RAM,
                                                               one to two minutes per case)
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "cases", "big_manifest.json")
REF = "/root/reference"

#        name     seed  record lengths              clean?  threads  id seed
CASES = {
    # one clean-path record, 1,124,799-ish candidates: slices [0,1e6) and the misplaced [r, 2r)
    "big1": (7, [9_000_000], True, 1, 31),
    # two formatted-path records: the second emission re-emits the first record (cumulative list,
    # CROPSR.py:407,442) and crosses 1e6 rows; 8 BLAS threads split the 1e6-row matmul
    "big2": (8, [5_000_000, 4_600_000], False, 8, 32),
}

RUNNER = r"""
import sys, runpy, time, os
import numpy as np
time.sleep = lambda s: None
np.random.seed(int(os.environ['GOLDEN_SEED']))
sys.path.insert(0, %r)
runpy.run_path(%r, run_name='__main__')
""" % (REF, os.path.join(REF, "CROPSR.py"))


def big_fasta(name):
    """FASTA text of a case: uniform i.i.d. ACGT (upper-case), numpy default_rng(seed)."""
    seed, lengths, clean, _, _ = CASES[name]
    rng = np.random.default_rng(seed)
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    recs = []
    for k, n in enumerate(lengths):
        body = lut[rng.integers(0, 4, size=n, dtype=np.uint8)].tobytes().decode("ascii")
        if clean:
            recs.append(f">chr{k + 1}\n{body}")
        else:
            recs.append(f">chr{k + 1}\n" + "\n".join(body[j:j + 80] for j in range(0, n, 80)))
    return "\n".join(recs) + ("" if clean else "\n")


def digests(csv_bytes):
    sys.path.insert(0, os.path.join(HERE, ".."))
    from helpers import normalised_digest
    text = csv_bytes.decode("utf-8")
    return {"csv_sha256": hashlib.sha256(csv_bytes).hexdigest(), "rows": csv_bytes.count(b"\r\n") - 1,
            "digest_no_id_no_score": normalised_digest(text), "digest_no_id_score_12g": normalised_digest(text, "%.12g")}


def run_case(name):
    seed, lengths, clean, threads, id_seed = CASES[name]
    with tempfile.TemporaryDirectory() as wd:
        fa = os.path.join(wd, name + ".fa")
        with open(fa, "w", newline="") as f:
            f.write(big_fasta(name))
        gff = os.path.join(HERE, "fixtures", "sample_genome.gff")
        env = dict(os.environ, GOLDEN_SEED=str(id_seed), OPENBLAS_NUM_THREADS=str(threads))
        p = subprocess.run([sys.executable, "-c", RUNNER, "-f", fa, "-g", gff, "-o", os.path.join(wd, "out.csv"), "--cas9"],
                           cwd=wd, env=env, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError(f"{name}: reference failed\n{p.stderr}")
        with open(os.path.join(wd, "out.csv"), "rb") as f:
            data = f.read()
    out = digests(data)
    out.update({"fasta_seed": seed, "lengths": lengths, "clean_path": clean, "blas_threads": threads, "seed": id_seed,
                "fasta_sha256": hashlib.sha256(big_fasta(name).encode()).hexdigest()})
    return out


def main():
    only = set(sys.argv[1:]) or set(CASES)
    manifest = json.load(open(OUT)) if os.path.exists(OUT) else {"cases": {}}
    for name in CASES:
        if name in only:
            manifest["cases"][name] = run_case(name)
            print(name, manifest["cases"][name]["rows"], "rows", flush=True)
    manifest["environment"] = {"numpy": np.__version__, "python": sys.version.split()[0],
                               "note": "unmodified /root/reference, time.sleep stubbed, ids seeded, AVX-512 host"}
    json.dump(manifest, open(OUT, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
