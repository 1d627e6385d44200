"""Drop-in for the two cropsr_functions entry points CROPSR.py calls
(/root/reference/CROPSR.py:66,70)."""
from cropsr_b200.ingest import formatted, generate_dictionary  # noqa: F401
