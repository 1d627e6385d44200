#!/usr/bin/env python3
"""Deterministically (re)generate the small FASTA fixtures under
tests/golden/fixtures/.  The fixtures are committed; this script documents how
they were made.  sample_genome.fa/.gff are the reference's own sample data
(yeast chrI, /root/reference/sample_data/), copied unchanged.
"""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
FIX = os.path.join(HERE, "fixtures")


def random_bases(rng, n, gc=0.45, lower_frac=0.15, n_frac=0.002, iupac_frac=0.001):
    p = [(1 - gc) / 2, (1 - gc) / 2, gc / 2, gc / 2]
    s = rng.choice(np.frombuffer(b"ATCG", dtype=np.uint8), size=n, p=p)
    # lowercase blocks
    i = 0
    while i < n:
        blk = int(rng.integers(20, 200))
        if rng.random() < lower_frac:
            s[i:i + blk] |= 0x20
        i += blk
    for frac, alphabet in ((n_frac, b"N"), (iupac_frac, b"RYKMSWnryUZuz")):
        k = int(n * frac)
        if k:
            idx = rng.choice(n, size=k, replace=False)
            s[idx] = rng.choice(np.frombuffer(alphabet, dtype=np.uint8), size=k)
    return s.tobytes().decode("ascii")


def wrap(seq, width):
    return "\n".join(seq[i:i + width] for i in range(0, len(seq), width))


def write(name, text):
    with open(os.path.join(FIX, name), "w", newline="") as f:
        f.write(text)


def main():
    os.makedirs(FIX, exist_ok=True)
    rng = np.random.default_rng(20240611)

    # 3 records, multi-line, trailing newline -> formatted path, cumulative emission
    recs = [("chrA", random_bases(rng, 1500)), ("chrB", random_bases(rng, 900, gc=0.6)),
            ("scaffold_3", random_bases(rng, 700, gc=0.3))]
    write("multi3.fa", "".join(f">{h}\n{wrap(s, 60)}\n" for h, s in recs))

    # 3 records, exactly two lines each, NO trailing newline -> clean path
    recs = [("c1", random_bases(rng, 1200)), ("c2", random_bases(rng, 600)),
            ("c3", random_bases(rng, 400, lower_frac=0.5))]
    write("clean3.fa", "\n".join(f">{h}\n{s}" for h, s in recs))

    # same shape but with a trailing newline -> formatted path
    write("clean3_trailing_nl.fa", "\n".join(f">{h}\n{s}" for h, s in recs) + "\n")

    # edge cases: PAM at position 0/1, CC/GG at the very end, lowercase PAMs,
    # N / IUPAC / U / Z inside windows, records shorter than one window,
    # an empty record, windows truncated at the token end on both strands.
    body = random_bases(rng, 300, lower_frac=0.0, n_frac=0, iupac_frac=0)
    e1 = "AGGCCTGGACC" + body + "CCAGGNCCGGTTCCNGGUCCZGGccaggCCtggAACCGG"
    e2 = "CCGG" * 20                      # dense overlapping hits, 80 bases
    e3 = "ACGTACGTAGGCC"                  # shorter than any window
    e4 = "GG" + random_bases(rng, 64, lower_frac=0, n_frac=0, iupac_frac=0) + "CC"
    e5 = "TTTTTTTTTTTTTTTTTTTTTTTTTAGGTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTT"   # exactly one + hit
    edge = [("e1", e1), ("e2", e2), ("e3", e3), ("e4", e4), ("e5", e5)]
    write("edge_clean.fa", "\n".join(f">{h}\n{s}" for h, s in edge))
    write("edge_fmt.fa", "".join(f">{h}\n{wrap(s, 50)}\n" for h, s in edge))

    # a single record with exactly one candidate: np.matmul with n == 1 (ddot order)
    write("single_candidate.fa", f">one\n{e5}")

    # headers with whitespace desynchronise the key/value pairing; duplicate keys
    write("ws_header.fa", f">chr1 some description\n{wrap(random_bases(rng, 400), 70)}\n"
                          f">chr2\n{wrap(random_bases(rng, 300), 70)}\n")
    d = random_bases(rng, 250)
    write("dup_keys.fa", f">dup\n{d}\n>other\n{random_bases(rng, 200)}\n>dup\n{random_bases(rng, 260)}")

    # an empty sequence line and a header-only record
    write("empty_records.fa", f">a\n\n>b\n{random_bases(rng, 120)}\n>c\n")

    # larger: >= 3840 candidates so that OpenBLAS gemv goes multi-threaded
    write("mid50k.fa", f">mid\n{wrap(random_bases(rng, 50000, lower_frac=0.05), 80)}\n")


if __name__ == "__main__":
    main()
