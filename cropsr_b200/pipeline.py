"""End-to-end --cas9 run: ingest -> GPU scan+score -> reference-order emission.

Mirror of /root/reference/CROPSR.py:333-486 main().  Differences that are
deliberate and documented in DESIGN.md: the whole genome is scanned on the GPU
in one launch sequence before the per-token emission loop starts, and the
reference's 5-second sleep per token (CROPSR.py:478) is not performed.
"""
import time

import numpy as np

from . import emit, engine, ingest


# A genome handle addresses its positions and its candidate rows with 32 bits (crp_genome_commit
# answers CRP_ERR_RANGE beyond that); a run keeps every handle well below, so that a 10 Gbp genome
# -- which the reference handles given enough RAM -- is simply several handles scanned in turn.
MAX_POSITIONS_PER_GENOME = 1 << 31


class HostRescorer:
    """x of a few candidates re-summed in another BLAS class from the HOST copy of the tokens:
    the 30 scored bytes of the window (CROPSR.py:458, long.replace('U','T').upper()) go through
    crp_rs1_preactivation.  Same values as crp_rescore (which reads the packed records of one
    genome handle); used where the rows of a slice span several handles or several GPUs."""

    def __init__(self, token_bytes):
        self.token_bytes = token_bytes

    def rescore(self, tok_index, t, strand, cls):
        rows = np.zeros((len(t), 30), dtype=np.uint8)
        for i, (k, ti, st) in enumerate(zip(np.asarray(tok_index).tolist(), np.asarray(t).tolist(),
                                            np.asarray(strand).tolist())):
            long_ = emit.guide_strings(self.token_bytes[k], int(ti), st in (b"-", "-"), 20)[1]
            b = long_.replace("U", "T").upper().encode("ascii")
            rows[i, :len(b)] = np.frombuffer(b, dtype=np.uint8)[:30]
        return engine.rs1_preactivation(rows, cls)


class ScanOutput:
    """Everything the emission loop needs from a finished scan, whichever way it ran (one
    handle, several handles on one GPU, several GPUs): per-token candidate rows in reference
    order, the counts, a rescorer, timings."""

    def __init__(self, token_bytes, parts, guide_len):
        self.token_bytes = token_bytes
        self.parts = parts                      # [(genome, result, first token, n tokens)]
        self.guide_len = guide_len
        n = len(token_bytes)
        self.seg_plus = np.zeros(n, dtype=np.uint64)
        self.seg_minus = np.zeros(n, dtype=np.uint64)
        self._rows = [None] * len(parts)
        for g, r, first, cnt in parts:
            self.seg_plus[first:first + cnt] = r.seg_plus
            self.seg_minus[first:first + cnt] = r.seg_minus
        self._part_of = np.zeros(n, dtype=np.int64)
        for i, (_, _, first, cnt) in enumerate(parts):
            self._part_of[first:first + cnt] = i
        self._host = HostRescorer(token_bytes)

    def locate(self, k):
        """(genome, result, segment index) of token k"""
        g, r, first, _ = self.parts[self._part_of[k]]
        return g, r, k - first

    def token_rows(self, k):
        """(t_plus, x_plus, t_minus, x_minus) of token k; the strands of a handle come over in one
        device-to-host copy each the first time one of its tokens is asked for."""
        i = int(self._part_of[k])
        g, r, first, _ = self.parts[i]
        if self._rows[i] is None:
            want = ("pos", "x")
            self._rows[i] = (r.fetch("+", want=want), r.fetch("-", want=want))
        plus, minus = self._rows[i]
        s = k - first
        a, b = int(r.off_plus[s]), int(r.off_plus[s + 1])
        c, d = int(r.off_minus[s]), int(r.off_minus[s + 1])
        sl = lambda col, lo, hi: None if col is None else col[lo:hi]
        return plus["pos"][a:b], sl(plus["x"], a, b), minus["pos"][c:d], sl(minus["x"], c, d)

    def rescore(self, tok_index, t, strand, cls):
        if len(self.parts) == 1:                # one handle: token index == segment index
            return self.parts[0][0].rescore(tok_index, t, strand, cls)
        return self._host.rescore(tok_index, t, strand, cls)

    def scan_ms(self):
        return float(sum(r.scan_ms() for _, r, _, _ in self.parts))

    def timing(self):
        t = {"h2d_ms": 0.0, "pack_ms": 0.0}
        for g, _, _, _ in self.parts:
            for k, v in g.timing().items():
                t[k] += v
        return t

    def free(self):
        for g, r, _, _ in self.parts:
            r.free()
            g.free()
        self.parts = []
        self._rows = []


def _groups(lengths, limit=None):
    """consecutive runs of tokens whose summed length stays below the per-handle limit"""
    limit = MAX_POSITIONS_PER_GENOME if limit is None else limit
    out, first, acc = [], 0, 0
    for k, n in enumerate(lengths):
        if k > first and acc + n > limit:
            out.append((first, k - first))
            first, acc = k, 0
        acc += n
    out.append((first, len(lengths) - first))
    return [g for g in out if g[1] > 0] or [(0, 0)]


def scan_tokens(tokens, guide_len=20, flags=0):
    """Pack every token of the ingest dict into HBM and scan it.
    Returns (genome, result, token_bytes) for callers that want the raw handles of a
    single-handle scan (tests, smoke); run_cas9 goes through scan_token_bytes."""
    genome = engine.Genome()
    token_bytes = []
    for value in tokens.values():
        b = value.encode("ascii") if isinstance(value, str) else bytes(value)
        token_bytes.append(b)
        genome.add_token(b)
    genome.commit()
    result = genome.scan(guide_len, flags)
    return genome, result, token_bytes


def scan_token_bytes(token_bytes, guide_len=20, flags=0, limit=None):
    """-> ScanOutput; tokens are spread over as many genome handles as the 32-bit limits ask for."""
    parts = []
    for first, cnt in _groups([len(b) for b in token_bytes], limit):
        genome = engine.Genome()
        for b in token_bytes[first:first + cnt]:
            genome.add_token(b)
        genome.commit()
        parts.append((genome, genome.scan(guide_len, flags), first, cnt))
    return ScanOutput(token_bytes, parts, guide_len)


def scan_fasta_file(fasta, guide_len=20, flags=0, limit=None):
    """Device-side ingest of a plain multi-line FASTA file: the file's bytes go to the GPU as they
    are, k_fasta_strip builds the tokens there, and the host gets them back for row formatting.
    Returns (keys, ScanOutput), or None when the file needs the literal host
    ingest (clean path, blanks in headers, duplicate names, ragged lines, ...)."""
    from ._native import CropsrError
    with open(fasta, "rb") as f:
        data = f.read()
    layout = ingest.plain_fasta_layout(data)
    if layout is None:
        return None
    arr = np.frombuffer(data, dtype=np.uint8)
    parts, token_bytes = [], []
    try:
        for first, cnt in _groups([rec[2] for rec in layout], limit):
            genome = engine.Genome()
            parts.append((genome, None, first, cnt))
            for key, off, nbytes, width, last in layout[first:first + cnt]:
                genome.add_fasta_record(arr, off, nbytes, width, last)
            genome.commit()
            parts[-1] = (genome, genome.scan(guide_len, flags), first, cnt)
            token_bytes += [genome.fetch_token(seg).tobytes() for seg in range(cnt)]
            genome.release_tokens()
    except CropsrError as e:
        for g, r, _, _ in parts:
            if r is not None:
                r.free()
            g.free()
        if e.code == -6:            # CRP_ERR_FORMAT: not plain after all
            return None
        raise
    return [rec[0] for rec in layout], ScanOutput(token_bytes, parts, guide_len)


def write_side_output(path, tokens, scan, gff_frame, flank, formatted_path, annotation_info=None):
    """Opt-in table of the per-candidate side outputs (one row per unique candidate, reference
    order).  NOT part of the reference's CSV: GC, poly-T / homopolymer flags, cut site, the
    +-L flank window, the GFF feature under the cut site (device kernels k_extras / k_annotate;
    semantics in oracle/extras_oracle.py) and the prmrdsgn2-style primer enumeration on the flank
    (k_primers, cropsr_b200/primers.py: primers passing the GC / Tm filter on either side, Tm-compatible
    pairs, the first pair; empty where the flank is shorter than e + l = 130 bases)."""
    import csv
    import pandas as pd
    from . import annotate, primers
    ivs = annotate.intervals_for_tokens(gff_frame, list(tokens.keys()), formatted_path)
    info = annotate.read_annotation_info(annotation_info) if annotation_info else {}    # -p: Phytozome annotation_info.txt
    columns = ["chromosome", "strand", "pam_pos", "cutsite", "gc", "poly_t", "homopolymer", "low_gc",
               "unscored_base", "longest_run", "flank_start", "flank_end", "feature_type", "feature_attributes",
               "fwd_primers", "rev_primers", "primer_pairs", "first_pair", "annotation_info"]
    with open(path, "w", newline="") as f:
        csv.writer(f, delimiter="\t").writerow(columns)
    # GFF rows are looked up per DISTINCT feature, then spread over the candidates: whole columns,
    # no Python work per candidate.  The device is asked once per strand of a genome handle -- every
    # segment's extras / annotation / primer windows in one call each -- so 20,000 scaffolds cost a
    # handful of host round trips; the rows are then put back into the reference's order
    # (token by token, '+' then '-').
    gff_feature = gff_frame["feature"].to_numpy(dtype=object)
    gff_attr = gff_frame["attributes"].to_numpy(dtype=object)
    keys = list(tokens.keys())
    for genome, result, first, cnt in scan.parts:
        part_ivs = ivs[first:first + cnt]
        iv_off = np.concatenate(([0], np.cumsum([len(iv["start"]) for iv in part_ivs]))).astype(np.uint64)
        iv_start = np.concatenate([iv["start"] for iv in part_ivs]) if cnt else np.empty(0, np.uint32)
        iv_end = np.concatenate([iv["end"] for iv in part_ivs]) if cnt else np.empty(0, np.uint32)
        iv_row = np.concatenate([np.asarray(iv["row"], dtype=np.int64) for iv in part_ivs]) if cnt else np.empty(0, np.int64)
        cols = {}
        for strand, seg_cnt in (("+", result.seg_plus), ("-", result.seg_minus)):
            seg_cnt = seg_cnt.astype(np.int64)
            n = int(seg_cnt.sum())
            seg_of = np.repeat(np.arange(cnt, dtype=np.int64), seg_cnt)
            pos = result.fetch(strand, want=("pos",))["pos"]
            ex = result.extras_strand(strand, flank)
            feat = result.annotate_strand(strand, iv_off, iv_start, iv_end)
            pr = primers.design_windows(genome, seg_of.astype(np.uint32), ex["flank_lo"], ex["flank_hi"])
            if n and len(iv_row):
                at = np.minimum(np.maximum(feat, 0) + iv_off[seg_of].astype(np.int64), len(iv_row) - 1)
                rows = np.where(feat >= 0, iv_row[at], -1)
            else:
                rows = np.full(n, -1, dtype=np.int64)
            ft = np.full(n, "", dtype=object)
            fa = np.full(n, "", dtype=object)
            ai = np.full(n, "", dtype=object)
            hit = rows >= 0
            if hit.any():
                uniq, inv = np.unique(rows[hit], return_inverse=True)
                ft[hit] = gff_feature[uniq][inv]
                fa[hit] = gff_attr[uniq][inv]
                ai[hit] = np.array([annotate.lookup_annotation_info(info, a) for a in gff_attr[uniq]], dtype=object)[inv]
            fl = ex["flags"].astype(np.int64)
            ok = pr["status"] == 0
            firstp = pr["first"].astype(np.int64)
            has_pair = ok & (firstp[:, 0] != 0xFFFF) if n else np.zeros(0, bool)
            pair = np.full(n, "", dtype=object)
            if has_pair.any():
                pair[has_pair] = [f"{a}+{b}/{c}+{d}" for a, b, c, d in firstp[has_pair].tolist()]
            blank_unless_ok = lambda a: np.where(ok, a.astype(np.int64).astype(str).astype(object), "")
            chrom = np.array([k[1:] for k in keys[first:first + cnt]], dtype=object)[seg_of] if n else np.empty(0, object)
            cols[strand] = (seg_of, pd.DataFrame({
                "chromosome": chrom, "strand": strand, "pam_pos": pos.astype(np.int64), "cutsite": ex["cut"].astype(np.int64),
                "gc": ex["gc"].astype(np.int64), "poly_t": fl & 1, "homopolymer": (fl >> 1) & 1, "low_gc": (fl >> 2) & 1,
                "unscored_base": (fl >> 3) & 1, "longest_run": ex["run"].astype(np.int64),
                "flank_start": ex["flank_lo"].astype(np.int64), "flank_end": ex["flank_hi"].astype(np.int64),
                "feature_type": ft, "feature_attributes": fa, "fwd_primers": blank_unless_ok(pr["n_fwd"]),
                "rev_primers": blank_unless_ok(pr["n_rev"]), "primer_pairs": blank_unless_ok(pr["n_pairs"]),
                "first_pair": pair, "annotation_info": ai}, columns=columns))
        # reference order: per token its '+' rows, then its '-' rows -- a stable sort on (token, strand)
        frame = pd.concat([cols["+"][1], cols["-"][1]], ignore_index=True)
        order = np.argsort(np.concatenate((2 * cols["+"][0], 2 * cols["-"][0] + 1)), kind="stable")
        frame.iloc[order].to_csv(path, sep="\t", mode="a", header=False, index=False, lineterminator="\r\n")


def write_gap_table(path, tokens, scan, min_len=10):
    """Opt-in companion of the side output: the gaps of the assembly -- runs of at least
    min_len bytes that are not ACGTacgt (N, IUPAC) -- one row per run, token coordinates, read off the packed
    records on the device (Genome.other_runs / k_other_runs).  Not part of the reference's output."""
    import csv
    with open(path, "w", newline="") as f:
        w = csv.writer(f, delimiter="\t")
        w.writerow(["chromosome", "start", "length"])
        for k, key in enumerate(tokens.keys()):
            genome, _, seg = scan.locate(k)
            start, length = genome.other_runs(seg, min_len)
            w.writerows((key[1:], int(a), int(b)) for a, b in zip(start.tolist(), length.tolist()))


def run_cas9(fasta, gff, output="data.csv", guide_len=20, verbose=False, blas_threads=1,
             time_path="time.txt", out=print, side_output=None, flank=200, device_ingest=True, annotation_info=None,
             chunk_rows=None, devices=None, handle_limit=None):
    """devices: CUDA ordinals; more than one shards the genome over one process per GPU
    (cropsr_b200/multi.py) and writes the same CSV.  handle_limit: positions per genome handle
    (tests shrink it to walk the several-handles path on small inputs)."""
    begin = time.time()
    timing = open(time_path, "w")                       # CROPSR.py:371
    multi_gpu = devices is not None and len(devices) > 1
    fast = scan_fasta_file(fasta, guide_len, limit=handle_limit) if device_ingest and not multi_gpu else None
    if fast is not None:                                # :374, on the device
        keys, scan = fast
        tokens = {k: b[:25].decode("ascii") for k, b in zip(keys, scan.token_bytes)}   # stdout only
        if verbose:
            out(f"Genome file {fasta} successfully imported")
            out("formatting genome")
            out(f"Genome file {fasta} successfully formatted")
            out("The genome was successfully converted to a dictionary")
    else:
        tokens = ingest.import_fasta_file(fasta, verbose)   # :374
    gff_frame = ingest.import_gff_file(gff, verbose)    # :375 (parsed, never used by the reference)
    if verbose:
        out("\n            Initiating PAM site detection.\n            \n"
            "            Please wait, this may take a while...\n            ")
    emit.write_header(output)                           # :402-405

    if fast is None:
        token_bytes = [v.encode("ascii") if isinstance(v, str) else bytes(v) for v in tokens.values()]
        if multi_gpu:
            from . import multi
            scan = multi.scan_on_devices(token_bytes, devices, guide_len)
        else:
            scan = scan_token_bytes(token_bytes, guide_len, limit=handle_limit)
    per_token = scan.seg_plus.astype(np.int64) + scan.seg_minus.astype(np.int64)
    table = emit.CandidateTable(guide_len, capacity=int(per_token.sum()))
    rows_written = 0
    # the cumulative list sizes of every emission are known once the scan is through: their ids
    # (CROPSR.py:448) are drawn ahead, in the reference's order, while rows are being written
    id_stream = emit.IdStream(np.cumsum(per_token))
    for k, (key, value) in enumerate(tokens.items()):
        out("Searching on Chromosome: ", key[:25])       # :410-411
        out("With start of sequence: ", value[:25])
        t_plus, x_plus, t_minus, x_minus = scan.token_rows(k)
        table.append_token(key, scan.token_bytes[k], k, t_plus, x_plus, t_minus, x_minus)
        if verbose:
            n = len(t_plus) + len(t_minus)
            out(f"\n                {n:n} Cas9 PAM sites were found on {key[1:]}\n                ")
        rows_written += emit.emit_cumulative(output, table, scan, blas_threads, id_stream, chunk_rows)
        timing.write("Total runtime of the program is " + str(time.time() - begin))   # :476-477
    id_stream.close()
    timing.close()
    if side_output and guide_len == 20:
        if multi_gpu:
            raise NotImplementedError("--side-output needs the packed genome on one device: run it without --devices")
        with open(fasta, "r") as f:
            formatted_path = ingest.needs_formatting(f.read())
        write_side_output(side_output, tokens, scan, gff_frame, flank, formatted_path, annotation_info)
        write_gap_table(side_output + ".gaps.tsv", tokens, scan)
    stats = {"tokens": len(tokens), "candidates": len(table), "rows": rows_written,
             "scan_ms": scan.scan_ms(), **scan.timing()}
    t_plus = x_plus = t_minus = x_minus = None          # views into the scan's row arrays
    scan.free()
    if verbose:
        out(f"The output file has been generated at {output}")
    return stats
