"""CPU ORACLE for the primer enumeration of prmrdsgn2.py.  TEST INFRASTRUCTURE ONLY:
only tests/ may import it; the product path (cropsr_b200/primers.py -> crp_primer_windows -> k_primers)
never does.

Restates, with plain Python strings and floats, the part of /root/reference/prmrdsgn2.py that runs
without bowtie2:
    Primer.__calculate_GC / __calculate_Tm      prmrdsgn2.py:76-95
    create_reverse_complement                   prmrdsgn2.py:104-112
    get_primers                                 prmrdsgn2.py:115-124
    filter_primers                              prmrdsgn2.py:127-137
    pairing of forward x reverse by Tm          prmrdsgn2.py:260-266
Parity status: PINNED to tests/golden/primer_vectors.json, which tests/golden/make_primer_golden.py
wrote by calling those functions of the unmodified reference module in the build container.
"""
import itertools
import math


def gc_percentage(seq):                      # prmrdsgn2.py:76-80
    up = seq.upper()
    return 100 * (float(up.count("G") + up.count("C")) / len(seq))


def melting_temp(seq):                       # prmrdsgn2.py:87-95
    n = len(seq)
    up = seq.upper()
    if n < 13:
        return (up.count("A") + up.count("T")) * 2 + (up.count("C") + up.count("G")) * 4
    return 64.9 + 41 * (up.count("G") + up.count("C") - 16.4) / n


def reverse_complement(seq):                 # prmrdsgn2.py:104-112: upper-case ACGT only, the rest unchanged
    comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
    return "".join(reversed([comp.get(b, b) for b in seq]))


def get_primers(seq, e, s, l):               # prmrdsgn2.py:115-124 -> [(i, sequence)] in the reference's order
    return [(i, seq[i:j + 1]) for i in range(e) for j in range(i + s, i + l)]


def passes(seq, m, x, M, X):                 # prmrdsgn2.py:133-135
    gc, tm = gc_percentage(seq), melting_temp(seq)
    return not (gc < M or gc > X or tm < m or tm > x)


def design(fragment, e=100, s=20, l=30, m=50, x=65, M=35, X=65, D=0.5):
    """-> dict(fwd=[(i, n)], rev=[(i, n)], n_pairs, first=(fi, fn, ri, rn) or None); rev indices are on
    the reverse complement of the fragment, as in the reference."""
    fwd = [(i, p) for i, p in get_primers(fragment, e, s, l) if passes(p, m, x, M, X)]
    rev = [(i, p) for i, p in get_primers(reverse_complement(fragment), e, s, l) if passes(p, m, x, M, X)]
    n_pairs, first = 0, None
    for (fi, fp), (ri, rp) in itertools.product(fwd, rev):
        if math.isclose(melting_temp(fp), melting_temp(rp), abs_tol=D):
            n_pairs += 1
            if first is None:
                first = (fi, len(fp), ri, len(rp))
    return {"fwd": [(i, len(p)) for i, p in fwd], "rev": [(i, len(p)) for i, p in rev], "n_pairs": n_pairs,
            "first": first}


def design_fast(fragment, e=100, s=20, l=30, m=50, x=65, M=35, X=65, D=0.5):
    """Same counts as design() without materialising the pairs (pairs counted per Tm value):
    used by the tests on many windows; checked against design() on the golden fragments."""
    def side(seq):
        out = []
        for i, p in get_primers(seq, e, s, l):
            if passes(p, m, x, M, X):
                out.append((i, len(p), melting_temp(p)))
        return out
    fwd, rev = side(fragment), side(reverse_complement(fragment))
    tms = sorted(set(t for _, _, t in rev))
    by_tm = {t: sum(1 for _, _, u in rev if u == t) for t in tms}
    n_pairs, first = 0, None
    cache = {}
    for fi, fn, ft in fwd:
        if ft not in cache:
            cache[ft] = [t for t in tms if math.isclose(ft, t, abs_tol=D)]
        ok = cache[ft]
        c = sum(by_tm[t] for t in ok)
        n_pairs += c
        if c and first is None:
            okset = set(ok)
            ri, rn, _ = next(r for r in rev if r[2] in okset)
            first = (fi, fn, ri, rn)
    return {"n_fwd": len(fwd), "n_rev": len(rev), "n_pairs": n_pairs, "first": first}
