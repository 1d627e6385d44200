"""GFF features -> per-token sorted interval arrays for the annotation kernel.

Opt-in: the reference parses the GFF (/root/reference/CROPSR.py:77-95) and never
uses it (:375 is the only reference; the CSV `features` column is the literal ''),
so nothing here touches the parity CSV.  The intent stated in the reference's
README.md:71-74 is "functional annotation of the site"; the semantics chosen
here -- innermost feature of the selected types containing the cut site -- are
pinned by oracle/extras_oracle.py.

Coordinates: GFF is 1-based inclusive.  A token of the *formatted* ingest path
holds base i (1-based) at index i (a leading quote sits at index 0), a token of
the *clean* path at index i - 1.
"""
import numpy as np

DEFAULT_FEATURES = ("gene", "CDS")


def token_chromosome(key, formatted_path):
    """Sequence name of an ingest-dict key: '>chr1' (clean) or "[('chr1'," / "('chr2'," (formatted)."""
    if not formatted_path:
        return key[1:]
    return key.strip("[](),").strip("'\"")


def intervals_for_tokens(frame, keys, formatted_path, features=DEFAULT_FEATURES):
    """frame: the DataFrame of ingest.import_gff_file.  Returns one dict per ingest key:
    start, end (uint32, inclusive token coordinates, sorted by start, longer feature first), row (index of
    the GFF row in `frame`), so that a feature index from the kernel maps back to its attributes."""
    shift = 0 if formatted_path else -1
    sel = frame[frame["feature"].isin(features)]
    sel = sel[np.isfinite(sel["start"].astype(float)) & np.isfinite(sel["end"].astype(float))]
    by_chrom = {c: g for c, g in sel.groupby("chromosome", sort=False)}
    out = []
    for key in keys:
        g = by_chrom.get(token_chromosome(key, formatted_path))
        if g is None or not len(g):
            out.append({"start": np.empty(0, np.uint32), "end": np.empty(0, np.uint32), "row": np.empty(0, np.int64)})
            continue
        start = g["start"].to_numpy(dtype=np.int64) + shift
        end = g["end"].to_numpy(dtype=np.int64) + shift
        order = np.lexsort((-end, start))      # equal starts: the shorter feature last, so that it wins
        out.append({"start": start[order].astype(np.uint32), "end": end[order].astype(np.uint32),
                    "row": g.index.to_numpy()[order]})
    return out
