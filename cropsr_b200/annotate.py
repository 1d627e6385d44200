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


# ------------------------------------------------------------------ Phytozome annotation_info.txt
# The reference accepts -p / --phytozome and only echoes the path (/root/reference/CROPSR.py:32,364);
# its README.md:71-74 states the intent ("functional annotation").  Here the file enriches the opt-in
# side output: the annotation_info row of the locus / transcript a GFF feature belongs to.  Semantics
# are this module's own (unpinned by the reference), tested in tests/test_host_logic.py.
PHYTOZOME_COLUMNS = ("pacId", "locusName", "transcriptName", "peptideName", "Pfam", "Panther", "KOG", "ec", "KO", "GO",
                     "Best-hit-arabi-name", "arabi-symbol", "arabi-defline")


def read_annotation_info(path):
    """Phytozome `*.annotation_info.txt` (tab separated, optional '#pacId ...' header line) ->
    dict: pacId / locusName / transcriptName / peptideName -> 'column=value;...' of the non-empty columns."""
    table = {}
    names = list(PHYTOZOME_COLUMNS)
    with open(path, "r") as f:
        for line in f:
            line = line.rstrip("\r\n")
            if not line:
                continue
            cells = line.split("\t")
            if line.startswith("#"):
                head = [c.lstrip("#").strip() for c in cells]
                if head and head[0] == "pacId":
                    names = head
                continue
            text = ";".join(f"{n}={c}" for n, c in zip(names[4:], cells[4:]) if c.strip())
            for key in cells[:4]:
                if key.strip():
                    table.setdefault(key.strip(), text)
    return table


def lookup_annotation_info(table, attributes):
    """GFF3 attribute string -> annotation_info text of the feature: the first of its pacid / ID / Name /
    Parent values (version suffixes like '.v3.1', feature suffixes like '.CDS.2' peeled off) the table knows."""
    if not table or not isinstance(attributes, str):
        return ""
    fields = dict(kv.split("=", 1) for kv in attributes.split(";") if "=" in kv)
    for name in ("pacid", "ID", "Name", "Parent"):
        for value in fields.get(name, "").split(","):
            v = value.strip()
            while v:
                if v in table:
                    return table[v]
                if "." not in v:
                    break
                v = v.rsplit(".", 1)[0]
    return ""
