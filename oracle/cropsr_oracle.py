"""CPU ORACLE for the CROPSR --cas9 scan + score + emission path.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; the product
package cropsr_b200/ never does, and fails loudly without its CUDA library.

This is a literal, deliberately slow restatement of what the unmodified
reference computes, function by function, bug for bug:

  ingest      /root/reference/CROPSR.py:54-74,
              /root/reference/cropsr_functions.py:190-196 (generate_dictionary),
              :221-229 (formatted)
  PAM scan    /root/reference/CROPSR.py:98-104, regexes at :415 and :426
  windows     /root/reference/CROPSR.py:418-423 (+), :429-434 (-)
  transforms  /root/reference/CROPSR.py:116-121 (reverse complement),
              :124-129 (gRNA)
  RS1 score   /root/reference/CROPSR.py:285-313 with constants :161-283
  ids         /root/reference/CROPSR.py:316-318, :448-449
  emission    /root/reference/CROPSR.py:386-405 (header), :442-474 (chunks),
              :155-158 (cut site), :476-478 (time.txt)

Parity status: PINNED.  tests/test_oracle_golden.py checks this module against
CSVs produced by the unmodified reference run in the build container
(tests/golden/make_golden.py is the generating script) and against the digests
of the reference's shipped sample_data/output.csv recorded in SURVEY.md section 4.

Third-party arithmetic the reference leans on, which lives outside
/root/reference (numpy, version unpinned by the reference; 2.3.5 with
scipy-openblas 0.3.30 in the build container):
  * np.matmul -> OpenBLAS dgemv_t / ddot.  The summation order is restated in
    `lane_sums` / `row_classes` below; it was derived by experiment against
    np.matmul in the build container and is re-checked by the tests whenever
    numpy is importable (score_blas vs score_model).
  * np.exp -> numpy's own SIMD kernel (not correctly rounded).  The oracle
    calls np.exp, exactly as the reference does.
"""
import csv
import io
import re
import time

import numpy as np

from rs1_table import W1, W2, INTERCEPT, LOW_GC

HEADER = ["crispr_id", "crispr_sys", "sequence", "long_sequence", "chromosome",
          "start_pos", "end_pos", "cutsite", "strand", "on_site_score",
          "features", "status"]                     # CROPSR.py:386-399
CHUNK = 1000000                                      # CROPSR.py:453
ALPHANUM = np.array(list("ABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789"), dtype="|U1")  # :316


# ----------------------------------------------------------------- ingest ---
def formatted(text):
    """cropsr_functions.py:221-229: records split on '>', header = first line,
    remaining newlines deleted, then the *repr of the list of tuples*."""
    records = []
    for chunk in text.split(">"):
        if not chunk:
            continue
        parts = chunk.split("\n", 1)
        records.append(tuple(p.replace("\n", "") for p in parts))
    return str(records)


def generate_dictionary(text):
    """cropsr_functions.py:190-196: whitespace tokens paired (key, value);
    an odd trailing token pairs with ''. dict => last value wins, first
    insertion position kept."""
    toks = text.split()
    out = {}
    for i in range(0, len(toks), 2):
        out[toks[i]] = toks[i + 1] if i + 1 < len(toks) else ""
    return out


def needs_formatting(text):
    """CROPSR.py:62-63."""
    return 2 * text.count(">") != text.count("\n") + 1


def import_fasta_text(text):
    """CROPSR.py:54-74 on already-read text."""
    if needs_formatting(text):
        text = formatted(text)
    return generate_dictionary(text)


# ------------------------------------------------------------- transforms ---
def reverse_complement(s):
    """CROPSR.py:116-121 (replace chain incl. its U->T and Z->G side effects)."""
    return (s.replace("A", "U").replace("C", "Z").replace("G", "C")
             .replace("Z", "G").replace("T", "A").replace("U", "T"))[::-1]


def grna(s):
    """CROPSR.py:124-129."""
    return (s.replace("A", "U").replace("C", "Z").replace("G", "C")
             .replace("Z", "G").replace("T", "A"))[::-1]


_PLUS = re.compile(r"(?=.GG)")      # CROPSR.py:415
_MINUS = re.compile(r"(?=CC.)")     # CROPSR.py:426


def pam_hits(tok, guide_len=20):
    """Positions t that survive the bounds tests of CROPSR.py:419 / :430."""
    L = len(tok)
    plus, minus = [], []
    for m in _PLUS.finditer(tok):
        t = m.start()
        p0, p1 = t - guide_len, t
        if p0 >= 5 and p0 + 5 <= L + 10 and p1 >= 5 and p1 <= L + 10:
            plus.append(t)
    for m in _MINUS.finditer(tok):
        t = m.start()
        p0, p1 = t + 3, t + 3 + guide_len
        if p0 >= 5 and p0 + 5 <= L + 10 and p1 >= 5 and p1 <= L + 10:
            minus.append(t)
    return plus, minus


def candidates_for_token(key, tok, guide_len=20):
    """CROPSR.py:413-434: records [start,end,chrom,short,long,'cas9',strand],
    all + hits by ascending t, then all - hits by ascending t."""
    plus, minus = pam_hits(tok, guide_len)
    chrom = key[1:]
    out = []
    for t in plus:
        p0, p1 = t - guide_len, t
        out.append([p0, p1, chrom, grna(tok[p0:p1]), grna(tok[p0 - 5:p1 + 5]), "cas9", "+"])
    for t in minus:
        p0, p1 = t + 3, t + 3 + guide_len
        out.append([p1, p0, chrom, grna(reverse_complement(tok[p0:p1])),
                    grna(reverse_complement(tok[p0 - 5:p1 + 5])), "cas9", "-"])
    return out


# ---------------------------------------------------------------- scoring ---
def scored_bytes(long_seq):
    """CROPSR.py:458: the uint8[30] row handed to rs1_score (valid 30-mers)."""
    return np.frombuffer(long_seq.replace("U", "T").upper().encode("ascii"), dtype=np.uint8)


_CMP1 = np.array([65, 84, 67, 71] * 30, dtype=np.float64)            # :300
_CMP2A = np.array(([65] * 4 + [84] * 4 + [67] * 4 + [71] * 4) * 29, dtype=np.float64)  # :301
_CMP2B = np.array([65, 84, 67, 71] * 4 * 29, dtype=np.float64)        # :302


def indicator_matrices(seqs):
    """CROPSR.py:288-309: the 0/1 float64 matrices (n,120) and (n,464)."""
    seqs = np.asarray(seqs)
    m1 = (np.repeat(seqs, 4, axis=1) == _CMP1).astype(np.float64)
    a = np.repeat(seqs[:, 0:29], 16, axis=1) == _CMP2A
    b = np.repeat(seqs[:, 1:30], 16, axis=1) == _CMP2B
    m2 = np.logical_and(a, b).astype(np.float64)
    return m1, m2


def score_blas(seqs):
    """CROPSR.py:285-313 verbatim in effect: np.matmul + np.exp.  The result
    depends on the BLAS build and thread count of the machine it runs on."""
    m1, m2 = indicator_matrices(seqs)
    sf = np.matmul(m1, W1)
    ss = np.matmul(m2, W2)
    x = (sf + ss + INTERCEPT + LOW_GC) * -1
    return 1 / (1 + np.exp(x))


# Row classes of one np.matmul call (OpenBLAS 0.3.30, x86-64 AVX2/AVX-512
# dgemv_t micro-kernels): which summation order row i of an n-row call gets.
CANONICAL, PAIR, SINGLE = 0, 1, 2
_MT_THRESHOLD = 460800      # gemv goes multi-threaded when rows*cols >= this


def _thread_ranges(n, d, threads):
    if threads <= 1 or n * d < _MT_THRESHOLD:
        return [(0, n)]
    ranges, left, pos, k = [], n, 0, 0
    while left > 0:
        w = (left + threads - k - 1) // (threads - k)
        w = max(w, 4)
        w = min(w, left)
        ranges.append((pos, pos + w))
        pos += w
        left -= w
        k += 1
    return ranges


def row_classes(n, d, threads=1):
    """Class of every row of an (n,d)@(d,) np.matmul.  n==1 -> ddot (SINGLE).
    Otherwise each thread's row range is walked 4 rows at a time (CANONICAL);
    a remainder of 2 or 3 sends its first two rows through the 2-lane kernel
    (PAIR); a lone last row is CANONICAL again."""
    cls = np.zeros(n, dtype=np.int8)
    if n == 1:
        cls[0] = SINGLE
        return cls
    for a, b in _thread_ranges(n, d, threads):
        w = b - a
        if w % 4 in (2, 3):
            base = a + (w // 4) * 4
            cls[base:base + 2] = PAIR
    return cls


def _seq_sum(m, w, cols):
    acc = np.zeros(m.shape[0])
    for j in cols:
        acc = acc + m[:, j] * w[j]
    return acc


def lane_sums(m, w, cls):
    """m (n,d) 0/1 float64, w (d,), cls (n,) -> (n,) sums in the order the
    BLAS kernel of each row's class uses."""
    n, d = m.shape
    out = np.empty(n)
    nz = [j for j in range(d) if w[j] != 0.0]
    sel = cls == CANONICAL
    if sel.any():
        p = [_seq_sum(m[sel], w, [j for j in nz if j % 4 == k]) for k in range(4)]
        out[sel] = (p[0] + p[2]) + (p[1] + p[3])
    sel = cls == PAIR
    if sel.any():
        q = [_seq_sum(m[sel], w, [j for j in nz if j % 2 == k]) for k in range(2)]
        out[sel] = q[0] + q[1]
    sel = cls == SINGLE
    if sel.any():
        ms = m[sel]
        n32 = d & ~31
        lanes = [[_seq_sum(ms, w, [j for j in range(8 * a + l, n32, 32)]) for l in range(8)]
                 for a in range(4)]
        f = [[lanes[a][i] + lanes[a][i + 4] for i in range(4)] for a in range(4)]
        pos = n32
        if d & 16:
            for a in range(4):
                for i in range(4):
                    j = pos + 4 * a + i
                    f[a][i] = f[a][i] + ms[:, j] * w[j]
            pos += 16
        t = [((f[0][i] + f[1][i]) + f[2][i]) + f[3][i] for i in range(4)]
        dot = (t[0] + t[2]) + (t[1] + t[3])
        for j in range(pos, d):
            dot = dot + ms[:, j] * w[j]
        out[sel] = dot
    return out


def preactivation_model(seqs, threads=1, classes=None):
    """x = -(((A+B)+intercept)+low_gc) per row with the modelled BLAS order."""
    seqs = np.asarray(seqs)
    n = len(seqs)
    if n == 0:
        return np.empty(0)
    m1, m2 = indicator_matrices(seqs)
    c1 = row_classes(n, 120, threads) if classes is None else classes
    c2 = row_classes(n, 464, threads) if classes is None else classes
    a = lane_sums(m1, W1, c1)
    b = lane_sums(m2, W2, c2)
    return (a + b + INTERCEPT + LOW_GC) * -1


def score_model(seqs, threads=1):
    return 1 / (1 + np.exp(preactivation_model(seqs, threads)))


# --------------------------------------------------------------- emission ---
def emission_slices(size, chunk=None):
    """CROPSR.py:451-472 replayed literally: the (start, count) of every slice
    that gets scored and written for a cumulative list of `size` rows.
    chunk: the literal 1000000 of CROPSR.py:453 (tests shrink it)."""
    CHUNK = globals()["CHUNK"] if chunk is None else chunk
    out = []
    count = 0
    counter = 0
    for i in range(size):
        count += 1
        if (count == CHUNK and i < size - 1) or (count < CHUNK and i == size - 1):
            out.append((count * counter, count))
            count = 0
            counter += 1
    return out


def emission_slices_closed_form(size, chunk=None):
    CHUNK = globals()["CHUNK"] if chunk is None else chunk
    q, r = divmod(size, CHUNK)
    if r > 0:
        return [(CHUNK * j, CHUNK) for j in range(q)] + [(r * q, r)]
    return [(CHUNK * j, CHUNK) for j in range(max(q - 1, 0))]


def make_ids(size):
    """CROPSR.py:316-318,448-449 on numpy's global legacy RNG."""
    ids = np.random.choice(ALPHANUM, [size, 7])
    return ["".join(row) for row in ids.tolist()]


def rows_for_slice(dataset, ids, start, count, score="blas", threads=1):
    """CROPSR.py:456-469 for one slice."""
    lesser = dataset[start:start + count]
    seqs = np.array([scored_bytes(it[4]) if len(it[4]) == 30 else np.empty(30,) for it in lesser])
    if score == "blas":
        sc = score_blas(seqs)
    else:
        with np.errstate(all="ignore"):
            sc = score_model(seqs, threads)
    rows = []
    for k, it in enumerate(lesser):
        rid = ids[start - k - 1]
        if len(it[4]) == 30:
            rows.append((rid, it[5], it[3], it[4], it[2], it[0], it[1], it[1] - 3, it[6],
                         sc[k], "", "completed"))
        else:
            rows.append((rid, it[5], it[3], it[4], it[2], it[0], it[1], it[6], -1, "", "completed"))
    return rows


def run(fasta_text, out, guide_len=20, score="blas", threads=1, log=None, chunk=None):
    """CROPSR.py:333-486 main() on FASTA text, writing the CSV to the text
    stream `out` (must have been opened with newline='').  Returns per-token
    unique candidate lists.  The 5 s sleep per token (:478) is omitted."""
    tokens = import_fasta_text(fasta_text)
    w = csv.writer(out)
    w.writerow(HEADER)
    dataset = []
    per_token = []
    for key, tok in tokens.items():
        if log is not None:
            log("Searching on Chromosome: ", key[:25])
            log("With start of sequence: ", tok[:25])
        cands = candidates_for_token(key, tok, guide_len)
        per_token.append(cands)
        dataset.extend(cands)
        size = len(dataset)
        ids = make_ids(size)
        for start, count in emission_slices(size, chunk):
            w.writerows(rows_for_slice(dataset, ids, start, count, score, threads))
    return per_token


def run_to_string(fasta_text, guide_len=20, score="blas", threads=1, chunk=None):
    buf = io.StringIO(newline="")
    run(fasta_text, buf, guide_len, score, threads, chunk=chunk)
    return buf.getvalue()


def timed_reference_pass(fasta_text, guide_len=20):
    """One full scan+score+emit pass (CSV into memory) for the cpu_baseline
    legs of bench.py.  Returns (seconds, n_bases_scanned, n_rows)."""
    t0 = time.perf_counter()
    buf = io.StringIO(newline="")
    per_token = run(fasta_text, buf, guide_len, "blas")
    dt = time.perf_counter() - t0
    n_bases = sum(len(v) for v in import_fasta_text(fasta_text).values())
    return dt, n_bases, buf.getvalue().count("\n") - 1
