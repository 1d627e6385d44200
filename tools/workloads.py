"""Synthetic genomes of BASELINE.json's configs: bench.py and the at-size parity
tests build their inputs here, chromosome by chromosome, each from its own seeded stream, so a
rank materialises only the records of its shard and every machine with this numpy gets the
same bytes (digests of the candidate tables are committed under tests/golden/).

    arabidopsis  configs[1]  seed 2, 5 chr {34,22,26,21,32} Mbp, GC 36 %, 15 % lower-case in 1-50 kb blocks
    sorghum      configs[2]  seed 3, 10 chr ~{81..61} x 1.07 Mbp, GC 44 %, 60 % lower-case
    maize        configs[3]  seed 4, 10 chr ~{307..150} x 1.09 Mbp, GC 47 %, 50 % lower-case,
                             implanted repeat library (soft-masked copies of 64 elements, ~40 % of the
                             bases), N runs of 100 b - 100 kb every ~1-5 Mbp
    sugarcane    configs[4]  seed 5, 100 pseudo-chromosomes ~80 Mbp + 20,000 scaffolds of 10-500 kb
                             (log-uniform), ~10.5 Gbp, GC 45 %, 40 % lower-case, short N runs

Records are returned as the reference's *formatted-path* tokens (quote/paren decoration
included) -- what CROPSR.py scans for any multi-line FASTA -- or written as an
80-column FASTA file for CLI runs.  arabidopsis / sorghum keep round 1's generator bit for bit
(their numbers stay comparable); maize / sugarcane use the byte-LUT generator below (5x faster:
10 Gbp has to be affordable).
"""
import numpy as np

_MBP = 1_000_000


def _sugarcane_lengths():
    r = np.random.default_rng(5005)
    chrom = (80 * _MBP * (0.9 + 0.2 * r.random(100))).astype(np.int64)
    scaf = np.exp(r.uniform(np.log(10_000), np.log(500_000), size=20_000)).astype(np.int64)
    return [int(x) for x in chrom] + [int(x) for x in scaf]


WORKLOADS = {
    # name: seed, record lengths, GC, lower-case fraction, style
    "sample": dict(seed=1, lengths=[230218], gc=0.38, lower=0.13, style="r1"),
    "arabidopsis": dict(seed=2, lengths=[34 * _MBP, 22 * _MBP, 26 * _MBP, 21 * _MBP, 32 * _MBP], gc=0.36, lower=0.15, style="r1"),
    "sorghum": dict(seed=3, lengths=[int(x * 1.07e6) for x in (81, 78, 74, 69, 72, 62, 65, 63, 59, 61)], gc=0.44, lower=0.60,
                    style="r1"),
    "maize": dict(seed=4, lengths=[int(x * 1.09e6) for x in (307, 244, 235, 247, 223, 174, 182, 181, 159, 150)], gc=0.47,
                  lower=0.50, style="r2", repeats=0.40, n_every=(1 * _MBP, 5 * _MBP), n_len=(100, 100_000)),
    "sugarcane": dict(seed=5, lengths=None, gc=0.45, lower=0.40, style="r2", repeats=0.0, n_every=(2 * _MBP, 10 * _MBP),
                      n_len=(100, 5_000)),
}
CONFIG_NAME = {
    "sample": "configs[0] sample-scale synthetic",
    "arabidopsis": "configs[1] synthetic Arabidopsis-scale 135 Mbp x5 chr",
    "sorghum": "configs[2] synthetic Sorghum-scale 730 Mbp x10 chr",
    "maize": "configs[3] synthetic maize-scale 2.3 Gbp x10 chr, repeat library + N runs",
    "sugarcane": "configs[4] synthetic sugarcane-scale 10.5 Gbp, 100 chr + 20,000 scaffolds",
}


def lengths(name):
    w = WORKLOADS[name]
    if w["lengths"] is None:
        w["lengths"] = _sugarcane_lengths()
    return w["lengths"]


def token_lengths(name):
    """len of every formatted-path token: bases + 4 decoration bytes."""
    return [n + 4 for n in lengths(name)]


def _bases_r1(name, k):
    """round 1's generator: float32 uniforms through searchsorted, then lower-case blocks"""
    w = WORKLOADS[name]
    n, gc = lengths(name)[k], w["gc"]
    lut = np.frombuffer(b"ATCG", dtype=np.uint8)
    thr = np.cumsum([(1 - gc) / 2, (1 - gc) / 2, gc / 2, gc / 2])
    r = np.random.default_rng(w["seed"] * 1000 + k)
    u = r.random(n, dtype=np.float32)
    s = lut[np.searchsorted(thr, u, side="right").clip(0, 3)]
    i = 0
    while i < n:
        blk = int(r.integers(1000, 50000))
        if r.random() < w["lower"]:
            s[i:i + blk] |= 0x20
        i += blk
    return s


def _byte_lut(gc):
    """256-entry byte -> base table: GC to 1/256 resolution"""
    n_gc = int(round(gc * 256))
    n_c = n_gc // 2
    n_a = (256 - n_gc) // 2
    return np.frombuffer(b"A" * n_a + b"T" * (256 - n_gc - n_a) + b"C" * n_c + b"G" * (n_gc - n_c), dtype=np.uint8)


_REPEAT_LIB = {}


def _repeat_library(name):
    if name not in _REPEAT_LIB:
        w = WORKLOADS[name]
        r = np.random.default_rng(w["seed"] * 7919)
        lib = []
        for _ in range(64):
            n = int(np.exp(r.uniform(np.log(200), np.log(9000))))
            gc = float(r.uniform(0.35, 0.65))
            lib.append(_byte_lut(gc)[np.frombuffer(r.bytes(n), dtype=np.uint8)] | 0x20)      # repeats are soft-masked
        _REPEAT_LIB[name] = lib
    return _REPEAT_LIB[name]


def _bases_r2(name, k):
    w = WORKLOADS[name]
    n = lengths(name)[k]
    r = np.random.default_rng(w["seed"] * 100_000 + k)
    s = _byte_lut(w["gc"])[np.frombuffer(r.bytes(n), dtype=np.uint8)]
    # lower-case (soft-masked) blocks of 1-50 kb
    if n >= 1000:
        n_blk = n // 25_000 + 2
        edges = np.minimum(np.cumsum(r.integers(1000, 50_000, size=n_blk)), n)
        low = r.random(n_blk) < w["lower"]
        start = 0
        for e, lo in zip(edges.tolist(), low.tolist()):
            if lo and e > start:
                s[start:e] |= 0x20
            start = e
            if start >= n:
                break
    # implanted repeat library: copies (with a few point substitutions) until ~repeats of the bases are covered
    if w.get("repeats", 0) > 0 and n > 20_000:
        lib = _repeat_library(name)
        mean_len = sum(len(x) for x in lib) / len(lib)
        n_ins = int(w["repeats"] * n / mean_len)
        which = r.integers(0, len(lib), size=n_ins)
        where = r.integers(0, n - 10_000, size=n_ins)
        for j, p in zip(which.tolist(), where.tolist()):
            el = lib[j]
            s[p:p + len(el)] = el
        # ~2 % divergence inside the copies would need per-base work; a sparse genome-wide substitution
        # pass keeps the copies from being identical
        n_sub = n // 100
        s[r.integers(0, n, size=n_sub)] = _byte_lut(w["gc"])[np.frombuffer(r.bytes(n_sub), dtype=np.uint8)]
    # N runs
    if w.get("n_every") and n > w["n_every"][0]:
        p = int(r.integers(*w["n_every"]))
        while p < n:
            ln = int(np.exp(r.uniform(np.log(w["n_len"][0]), np.log(w["n_len"][1]))))
            s[p:p + ln] = ord("N")
            p += ln + int(r.integers(*w["n_every"]))
    return s


def bases(name, k):
    """uint8 array of the bases of record k"""
    return (_bases_r1 if WORKLOADS[name]["style"] == "r1" else _bases_r2)(name, k)


def token(name, k):
    """record k as the reference's formatted-path token: ' + bases + '), (or ')] for the last record)"""
    n_rec = len(lengths(name))
    tail = b"')," if k + 1 < n_rec else b"')]"
    return np.concatenate((np.frombuffer(b"'", np.uint8), bases(name, k), np.frombuffer(tail, np.uint8)))


def tokens(name, only=None):
    """list of tokens (None for records outside `only`)"""
    return [token(name, k) if only is None or k in only else None for k in range(len(lengths(name)))]


def record_name(name, k):
    n_chr = {"sugarcane": 100}.get(name, len(lengths(name)))
    return f"Chr{k + 1:02d}" if k < n_chr else f"scaffold_{k - n_chr + 1}"


def write_fasta(name, path, width=80, records=None, prefix_bases=None):
    """80-column multi-line FASTA with a trailing newline (-> the reference's formatted path).
    records: indices to write (default all); prefix_bases: only the first so many bases of each."""
    ks = range(len(lengths(name))) if records is None else records
    with open(path, "wb") as f:
        for k in ks:
            s = bases(name, k)
            if prefix_bases is not None:
                s = s[:prefix_bases]
            f.write(b">" + record_name(name, k).encode() + b"\n")
            n = len(s)
            full = n // width * width
            if full:
                body = np.empty((full // width, width + 1), dtype=np.uint8)
                body[:, :width] = s[:full].reshape(-1, width)
                body[:, width] = 10
                f.write(body.tobytes())
            if n > full:
                f.write(s[full:].tobytes() + b"\n")


def write_gff(name, path, records=None, prefix_bases=None):
    """Phytozome-style GFF3: genes every ~5-20 kb with mRNA / exon / CDS children, sorted."""
    ks = range(len(lengths(name))) if records is None else records
    r = np.random.default_rng(WORKLOADS[name]["seed"] * 31 + 7)
    with open(path, "w") as f:
        f.write("##gff-version 3\n")
        gid = 0
        for k in ks:
            n = lengths(name)[k] if prefix_bases is None else min(lengths(name)[k], prefix_bases)
            chrom = record_name(name, k)
            p = int(r.integers(1000, 20000))
            while p + 6000 < n:
                gl = int(r.integers(800, 5000))
                strand = "+" if r.random() < 0.5 else "-"
                gid += 1
                g = f"Synth.{gid:06d}"
                f.write(f"{chrom}\tsynth\tgene\t{p}\t{p + gl}\t.\t{strand}\t.\tID={g};Name={g}\n")
                f.write(f"{chrom}\tsynth\tmRNA\t{p}\t{p + gl}\t.\t{strand}\t.\tID={g}.1;Parent={g};pacid={gid}\n")
                e1 = p + gl // 3
                e2 = p + 2 * gl // 3
                for a, b, j in ((p, e1, 1), (e2, p + gl, 2)):
                    f.write(f"{chrom}\tsynth\texon\t{a}\t{b}\t.\t{strand}\t.\tID={g}.1.exon.{j};Parent={g}.1\n")
                    f.write(f"{chrom}\tsynth\tCDS\t{a}\t{b}\t.\t{strand}\t0\tID={g}.1.CDS.{j};Parent={g}.1\n")
                p += gl + int(r.integers(5000, 20000))
