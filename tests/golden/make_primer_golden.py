#!/usr/bin/env python3
"""Golden vectors for the primer enumeration, produced by the UNMODIFIED reference module
/root/reference/prmrdsgn2.py in the build container (its argparse runs at import, so sys.argv is
set first; nothing of it is patched).  Only the functions that need no bowtie2 are called:
get_primers, filter_primers, create_reverse_complement, the Primer class and the Tm pairing
expression of main() (prmrdsgn2.py:260-266).  Output: tests/golden/primer_vectors.json.

usage: python tests/golden/make_primer_golden.py
"""
import hashlib
import itertools
import json
import math
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference(argv):
    sys.argv = ["prmrdsgn2.py", "-i", "unused", "-g", "unused"] + argv
    sys.path.insert(0, "/root/reference")
    for k in [k for k in sys.modules if k == "prmrdsgn2"]:
        del sys.modules[k]
    import prmrdsgn2
    return prmrdsgn2


def fragments():
    rng = random.Random(20261018)
    out = []
    for n, gc, lower, junk in ((400, 0.5, 0.0, 0.0), (400, 0.36, 0.15, 0.0), (400, 0.62, 0.3, 0.01), (260, 0.45, 0.1, 0.02),
                               (130, 0.5, 0.0, 0.0), (401, 0.2, 0.0, 0.0), (400, 0.8, 0.5, 0.0), (1000, 0.47, 0.5, 0.003)):
        s = []
        for _ in range(n):
            r = rng.random()
            b = rng.choice("GC") if r < gc else rng.choice("AT")
            if rng.random() < lower:
                b = b.lower()
            if rng.random() < junk:
                b = rng.choice("NnRY'),")
            s.append(b)
        out.append("".join(s))
    out.append("G" * 200 + "C" * 200)
    out.append("ACGT" * 100)
    return out


def main():
    cases = []
    for argv in ([], ["-e", "64", "-s", "17", "-l", "24", "-m", "48", "-x", "62.5", "-M", "30", "-X", "70", "-D", "0.25"],
                 ["-e", "37", "-s", "24", "-l", "32", "-D", "2"]):
        ref = load_reference(argv)
        a = ref.args
        for frag in fragments():
            if len(frag) < a.e + a.l:
                continue
            fwd = ref.filter_primers(ref.get_primers(frag))
            rc = ref.create_reverse_complement(frag)
            rev = ref.filter_primers(ref.get_primers(rc))
            pairs = [p for p in itertools.product(fwd, rev) if math.isclose(p[0].Tm, p[1].Tm, abs_tol=a.D)]
            h = hashlib.sha256()
            for p in fwd + rev:
                h.update(repr((p.sequence, p.GC_percentage, p.Tm)).encode())
            first = None
            if pairs:
                # (i, n) of the first pair: the first enumerated primer with that content (an earlier primer
                # with the same sequence has the same GC / Tm, so it would have been the first itself)
                def index_of(target, seq):
                    for k, (i, j) in enumerate((i, j) for i in range(a.e) for j in range(i + a.s, i + a.l)):
                        if seq[i:j + 1] == target.sequence:
                            return i, j + 1 - i
                    raise AssertionError
                first = list(index_of(pairs[0][0], frag) + index_of(pairs[0][1], rc))
            cases.append({"params": {"e": a.e, "s": a.s, "l": a.l, "m": a.m, "x": a.x, "M": a.M, "X": a.X, "D": a.D},
                          "fragment": frag, "n_fwd": len(fwd), "n_rev": len(rev), "n_pairs": len(pairs),
                          "first": first, "sha256_passing": h.hexdigest()})
    with open(os.path.join(HERE, "primer_vectors.json"), "w") as f:
        json.dump({"source": "/root/reference/prmrdsgn2.py get_primers / filter_primers / create_reverse_complement / Primer, "
                             "pairing as prmrdsgn2.py:260-266", "cases": cases}, f, indent=0)
    print(f"{len(cases)} cases")


if __name__ == "__main__":
    main()
