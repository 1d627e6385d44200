import csv
import gzip
import hashlib
import io
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def golden_csv(name):
    with gzip.open(os.path.join(GOLDEN, "cases", name + ".csv.gz")) as f:
        return f.read().decode("utf-8")


def fixture_text(fasta):
    with open(os.path.join(GOLDEN, "fixtures", fasta), newline="") as f:
        return f.read()


def fixture_path(fasta):
    return os.path.join(GOLDEN, "fixtures", fasta)


def normalised_digest(csv_text, score_fmt=None):
    """sha256 over rows joined with \\x1f, newline-terminated,
    header included, read with csv.reader; crispr_id dropped; on_site_score
    dropped (score_fmt None) or reformatted ('%.12g')."""
    h = hashlib.sha256()
    for n, row in enumerate(csv.reader(io.StringIO(csv_text, newline=""))):
        row = row[1:]
        if len(row) == 11:          # normal row / header: score is column 9 of 12
            if score_fmt is None:
                row = row[:8] + row[9:]
            elif n > 0:
                row[8] = score_fmt % float(row[8])
        h.update(("\x1f".join(row) + "\n").encode())
    return h.hexdigest()


def synthetic_fasta(seed, lengths, gc=0.45, lower_frac=0.15, n_frac=0.001, width=80, trailing_newline=True):
    """Multi-record FASTA text: i.i.d. bases at the given GC, lower-case blocks, a few N."""
    rng = np.random.default_rng(seed)
    p = [(1 - gc) / 2, (1 - gc) / 2, gc / 2, gc / 2]
    out = []
    for k, n in enumerate(lengths):
        s = rng.choice(np.frombuffer(b"ATCG", dtype=np.uint8), size=n, p=p)
        i = 0
        while i < n:
            blk = int(rng.integers(1000, 50000))
            if rng.random() < lower_frac:
                s[i:i + blk] |= 0x20
            i += blk
        k_n = int(n * n_frac)
        if k_n:
            s[rng.choice(n, size=k_n, replace=False)] = ord("N")
        body = s.tobytes().decode("ascii")
        lines = "\n".join(body[j:j + width] for j in range(0, n, width))
        out.append(f">chr{k + 1}\n{lines}")
    return "\n".join(out) + ("\n" if trailing_newline else "")


# ---- Python mirrors of the reference's row tuples: what the library's C row formatter (crp_format_rows) is compared with
def ids_to_strings(ids):
    ids = np.ascontiguousarray(ids)
    if ids.size == 0:
        return []
    return ids.view("<U7").ravel().tolist()


def slice_rows(table, ids, scores, scored, start, count):
    """Row tuples of one emitted slice (CROPSR.py:463-469)."""
    from cropsr_b200 import emit
    l = table.guide_len
    rows = []
    tok_i = table.tok[start:start + count].tolist()
    minus = table.minus[start:start + count].tolist()
    ts = table.t[start:start + count].tolist()
    sc = scores.tolist()
    ok = scored.tolist()
    for k in range(count):
        t = ts[k]
        token = table.tokens[tok_i[k]]
        short, long_ = emit.guide_strings(token, t, minus[k], l)
        if minus[k]:
            first, second, strand = t + 3 + l, t + 3, "-"
        else:
            first, second, strand = t - l, t, "+"
        rid = ids[start - k - 1]
        chrom = table.chroms[tok_i[k]]
        if ok[k]:
            rows.append((rid, "cas9", short, long_, chrom, first, second,
                         emit.apply_cutsite(first, second, "cas9"), strand, sc[k], "", "completed"))
        else:
            rows.append((rid, "cas9", short, long_, chrom, first, second, strand, -1, "", "completed"))
    return rows


