#!/usr/bin/env python3
"""Drop-in for the reference's CROPSR.py command line: same flags
(-f/-g/-p/-o/-l/-L/--cas9/-v), same CSV, the scan and scoring run on a B200."""
import sys

from cropsr_b200.cli import main
from cropsr_b200.refapi import (alphanum, apply_cutsite, find_PAM_site, get_gRNA_sequence, get_id,  # noqa: F401
                                get_reverse_complement, import_fasta_file, import_gff_file, rs1_score)

if __name__ == "__main__":
    sys.exit(main())
