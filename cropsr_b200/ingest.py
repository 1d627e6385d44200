"""FASTA / GFF ingest with the reference's exact (quirky) semantics.

Host-side mirror of /root/reference/CROPSR.py:54-95 and of the two
cropsr_functions entry points it calls (/root/reference/cropsr_functions.py
:190-196 ``generate_dictionary``, :221-229 ``formatted``).  The values of the
returned dict are the *tokens* the GPU scans -- in the formatted path they
include the quote/paren decoration that the reference's ``str(list_of_tuples)``
round-trip leaves around every sequence.
"""


def formatted(input_genome):
    """Multi-line FASTA text -> the repr() of [(header, sequence), ...]."""
    records = []
    for record in input_genome.split(">"):
        if record == "":
            continue
        head_and_body = record.split("\n", 1)
        records.append(tuple(part.replace("\n", "") for part in head_and_body))
    return repr(records)


def generate_dictionary(input):
    """Whitespace-separated tokens paired up as {key: value}; an unpaired last
    token maps to ''.  Later duplicates overwrite the value, not the position."""
    tokens = input.split()
    if len(tokens) % 2:
        tokens.append("")
    return dict(zip(tokens[0::2], tokens[1::2]))


def needs_formatting(text):
    """The reference's "is this already two lines per record" test
    (CROPSR.py:62-63): false only for exactly 2 lines/record, no trailing newline."""
    return 2 * text.count(">") != text.count("\n") + 1


def fasta_text_to_tokens(text, verbose=False, name="<text>"):
    if needs_formatting(text):
        if verbose:
            print("formatting genome")
        text = formatted(text)
        if verbose:
            print(f"Genome file {name} successfully formatted")
    genome = generate_dictionary(text)
    if verbose:
        print("The genome was successfully converted to a dictionary")
    return genome


def import_fasta_file(fasta, verbose=False):
    """CROPSR.py:54-74."""
    with open(fasta, "r") as f:
        text = f.read()
    if verbose:
        print(f"Genome file {fasta} successfully imported")
    return fasta_text_to_tokens(text, verbose, fasta)


def plain_fasta_layout(buf):
    """Layout of a *plain* multi-line FASTA file for the device ingest
    (engine.Genome.add_fasta_record), or None if the literal text path above is needed.

    buf: bytes-like, the whole file.  Returns [(key, seq_off, seq_bytes, line_width, last)],
    key being the ingest-dict key the reference derives for the record ("[('name'," for the
    first record, "('name'," for the others -- CROPSR.py:66-70 via str(list_of_tuples).split()).
    Plain means: the formatted path applies (CROPSR.py:62-63), the file starts with '>', every
    '>' starts a line, header lines hold no blank, quote, backslash or non-ASCII byte, no two
    records share a name, and every record has a sequence part; the sequence bytes themselves
    are validated on the device (k_fasta_strip)."""
    data = bytes(buf) if not isinstance(buf, (bytes, bytearray)) else buf
    n = len(data)
    if n == 0 or data[0:1] != b">":
        return None
    n_gt, n_nl = data.count(b">"), data.count(b"\n")
    if 2 * n_gt == n_nl + 1:                     # the clean path: tokens come from text.split()
        return None
    if data.count(b"\n>") + 1 != n_gt or b"\r" in data:
        return None
    records, seen, pos = [], set(), 0
    while pos < n:
        eol = data.find(b"\n", pos)
        if eol < 0:
            return None                          # header without a sequence part
        name = data[pos + 1:eol]
        nxt = data.find(b">", eol + 1)
        end = n if nxt < 0 else nxt
        if end - (eol + 1) <= 0 or not name or name in seen:
            return None
        if any(c <= 0x20 or c >= 0x7F or c in b"'\"\\" for c in name):
            return None
        seen.add(name)
        first_nl = data.find(b"\n", eol + 1, end)
        width = (first_nl if first_nl >= 0 else end) - (eol + 1)
        if width <= 0:
            return None
        text = name.decode("ascii")
        key = ("[('" if not records else "('") + text + "',"
        records.append((key, eol + 1, end - (eol + 1), width, nxt < 0))
        pos = end
    return records


GFF_COLUMNS = ["chromosome", "source", "feature", "start", "end", "score", "strand", "phase", "attributes"]


def import_gff_file(gff, verbose=False):
    """CROPSR.py:77-95: skip the leading lines that contain '##', then read a
    9-column tab-separated frame.  ``gff=None`` raises TypeError exactly like
    the reference's ``open(None)``."""
    import pandas as pd
    start_index = 0
    with open(gff, "r") as raw:
        if verbose:
            print(f"Annotation file {gff} successfully imported")
        for index, line in enumerate(raw):
            if "##" not in line:
                start_index = index
                break
    frame = pd.read_csv(gff, sep="\t", skiprows=start_index, header=None, names=GFF_COLUMNS)
    if verbose:
        print("Annotation database successfully generated")
    return frame
