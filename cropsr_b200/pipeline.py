"""End-to-end --cas9 run: ingest -> GPU scan+score -> reference-order emission.

Mirror of /root/reference/CROPSR.py:333-486 main().  Differences that are
deliberate and documented in DESIGN.md: the whole genome is scanned on the GPU
in one launch sequence before the per-token emission loop starts, and the
reference's 5-second sleep per token (CROPSR.py:478) is not performed.
"""
import time

import numpy as np

from . import emit, engine, ingest


def scan_tokens(tokens, guide_len=20, flags=0):
    """Pack every token of the ingest dict into HBM and scan it.
    Returns (genome, result, token_bytes)."""
    genome = engine.Genome()
    token_bytes = []
    for value in tokens.values():
        b = value.encode("ascii") if isinstance(value, str) else bytes(value)
        token_bytes.append(b)
        genome.add_token(b)
    genome.commit()
    result = genome.scan(guide_len, flags)
    return genome, result, token_bytes


def scan_fasta_file(fasta, guide_len=20, flags=0):
    """Device-side ingest of a plain multi-line FASTA file: the file's bytes go to the GPU as they
    are, k_fasta_strip builds the tokens there, and the host gets them back for row formatting.
    Returns (keys, genome, result, token_bytes), or None when the file needs the literal host
    ingest (clean path, blanks in headers, duplicate names, ragged lines, ...)."""
    from ._native import CropsrError
    with open(fasta, "rb") as f:
        data = f.read()
    layout = ingest.plain_fasta_layout(data)
    if layout is None:
        return None
    arr = np.frombuffer(data, dtype=np.uint8)
    genome = engine.Genome()
    try:
        for key, off, nbytes, width, last in layout:
            genome.add_fasta_record(arr, off, nbytes, width, last)
        genome.commit()
    except CropsrError as e:
        genome.free()
        if e.code == -6:            # CRP_ERR_FORMAT: not plain after all
            return None
        raise
    result = genome.scan(guide_len, flags)
    token_bytes = [genome.fetch_token(seg).tobytes() for seg in range(len(layout))]
    genome.release_tokens()
    return [rec[0] for rec in layout], genome, result, token_bytes


def write_side_output(path, tokens, result, gff_frame, flank, formatted_path, annotation_info=None):
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
    # one strand of one token at a time, no Python work per candidate
    gff_feature = gff_frame["feature"].to_numpy(dtype=object)
    gff_attr = gff_frame["attributes"].to_numpy(dtype=object)
    for seg, key in enumerate(tokens.keys()):
        iv = ivs[seg]
        for strand in "+-":
            pos = result.fetch_segment(seg, strand, want=("pos",))["pos"]
            n = len(pos)
            if n == 0:
                continue
            ex = result.extras(seg, strand, flank)
            feat = result.annotate(seg, strand, iv["start"], iv["end"])
            pr = primers.design_windows(result.genome, np.full(n, seg, np.uint32), ex["flank_lo"], ex["flank_hi"])
            rows = np.where(feat >= 0, iv["row"][np.maximum(feat, 0)] if len(iv["row"]) else -1, -1)
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
            first = pr["first"].astype(np.int64)
            has_pair = ok & (first[:, 0] != 0xFFFF)
            pair = np.full(n, "", dtype=object)
            if has_pair.any():
                fp = first[has_pair]
                pair[has_pair] = [f"{a}+{b}/{c}+{d}" for a, b, c, d in fp.tolist()]
            blank_unless_ok = lambda a: np.where(ok, a.astype(np.int64).astype(str).astype(object), "")
            frame = pd.DataFrame({
                "chromosome": key[1:], "strand": strand, "pam_pos": pos.astype(np.int64), "cutsite": ex["cut"].astype(np.int64),
                "gc": ex["gc"].astype(np.int64), "poly_t": fl & 1, "homopolymer": (fl >> 1) & 1, "low_gc": (fl >> 2) & 1,
                "unscored_base": (fl >> 3) & 1, "longest_run": ex["run"].astype(np.int64),
                "flank_start": ex["flank_lo"].astype(np.int64), "flank_end": ex["flank_hi"].astype(np.int64),
                "feature_type": ft, "feature_attributes": fa, "fwd_primers": blank_unless_ok(pr["n_fwd"]),
                "rev_primers": blank_unless_ok(pr["n_rev"]), "primer_pairs": blank_unless_ok(pr["n_pairs"]),
                "first_pair": pair, "annotation_info": ai}, columns=columns)
            frame.to_csv(path, sep="\t", mode="a", header=False, index=False, lineterminator="\r\n")


def run_cas9(fasta, gff, output="data.csv", guide_len=20, verbose=False, blas_threads=1,
             time_path="time.txt", out=print, side_output=None, flank=200, device_ingest=True, annotation_info=None,
             chunk_rows=None):
    begin = time.time()
    timing = open(time_path, "w")                       # CROPSR.py:371
    fast = scan_fasta_file(fasta, guide_len) if device_ingest else None
    if fast is not None:                                # :374, on the device
        keys, genome, result, token_bytes = fast
        tokens = {k: b.decode("ascii") for k, b in zip(keys, (tb[:25] for tb in token_bytes))}   # stdout only
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
        genome, result, token_bytes = scan_tokens(tokens, guide_len)
    table = emit.CandidateTable(guide_len)
    rows_written = 0
    # the cumulative list sizes of every emission are known once the scan is through: their ids
    # (CROPSR.py:448) are drawn ahead, in the reference's order, while rows are being written
    id_stream = emit.IdStream(np.cumsum(result.seg_plus.astype(np.int64) + result.seg_minus.astype(np.int64)))
    for seg, (key, value) in enumerate(tokens.items()):
        out("Searching on Chromosome: ", key[:25])       # :410-411
        out("With start of sequence: ", value[:25])
        plus = result.fetch_segment(seg, "+", want=("pos", "x"))
        minus = result.fetch_segment(seg, "-", want=("pos", "x"))
        table.append_token(key, token_bytes[seg], seg, plus["pos"], plus["x"], minus["pos"], minus["x"])
        if verbose:
            n = len(plus["pos"]) + len(minus["pos"])
            out(f"\n                {n:n} Cas9 PAM sites were found on {key[1:]}\n                ")
        rows_written += emit.emit_cumulative(output, table, genome, blas_threads, id_stream, chunk_rows)
        timing.write("Total runtime of the program is " + str(time.time() - begin))   # :476-477
    id_stream.close()
    timing.close()
    if side_output and guide_len == 20:
        with open(fasta, "r") as f:
            formatted_path = ingest.needs_formatting(f.read())
        write_side_output(side_output, tokens, result, gff_frame, flank, formatted_path, annotation_info)
    stats = {"tokens": len(tokens), "candidates": len(table), "rows": rows_written,
             "scan_ms": result.scan_ms(), **genome.timing()}
    result.free()
    genome.free()
    if verbose:
        out(f"The output file has been generated at {output}")
    return stats
