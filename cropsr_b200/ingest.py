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
