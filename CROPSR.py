#!/usr/bin/env python3
"""Drop-in for the reference's CROPSR.py command line: same flags
(-f/-g/-p/-o/-l/-L/--cas9/-v), same CSV, the scan and scoring run on a B200."""
import sys

from cropsr_b200.cli import main

if __name__ == "__main__":
    sys.exit(main())
